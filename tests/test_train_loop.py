"""Host logic of g2048.ppo.train_loop.PPOIterationLoop (no GPU: the rollouts are faked): the stopping rule and
checkpoint cadence of PPOTrainer.train (src/ppo/ppo_trainer.py:625-728), the bounded episode history and its tail
statistics (the reference slices a deque there and raises), resume."""
from collections import deque

import numpy as np
import pytest

from g2048.ppo import train_loop as TL


class FakeBuffer:
    buffer_size = 0


class FakeRunner:
    def __init__(self):
        self.position = 0

    def state_dict(self):
        return {"position": self.position}

    def load_state_dict(self, state):
        self.position = state["position"]


@pytest.fixture
def fake_collect(monkeypatch):
    """Every batch: batch_size episodes of length 10 + (episode number % 5), reward = episode number."""
    def collect(runner, buffer, batch_size, num_batches, reset=True):
        n = batch_size * num_batches
        ids = np.arange(runner.position, runner.position + n)
        runner.position += n
        lengths = 10 + ids % 5
        buffer.buffer_size = int(lengths.sum())
        return {"episode_rewards": ids.astype(np.float32), "episode_lengths": lengths.astype(np.int64),
                "total_episodes": n, "timesteps": buffer.buffer_size}

    monkeypatch.setattr(TL, "_collect", collect)
    return collect


def test_tail_of_a_deque():
    d = deque(range(10), maxlen=10)
    with pytest.raises(TypeError):
        d[-3:]  # what ppo_trainer.py:237 does
    assert TL.tail(d, 3) == [7, 8, 9]
    assert TL.tail(d, 100) == list(range(10))
    assert TL.tail(d, 0) == [] and TL.tail(deque(), 5) == []


def test_exactly_one_learner():
    with pytest.raises(ValueError):
        TL.PPOIterationLoop(FakeRunner(), FakeBuffer())
    with pytest.raises(ValueError):
        TL.PPOIterationLoop(FakeRunner(), FakeBuffer(), minibatch_step=lambda b: {}, update_policy=lambda **k: {})


def test_train_stops_at_the_target_and_checkpoints_like_the_reference(fake_collect):
    updates, saved, said = [], [], []
    loop = TL.PPOIterationLoop(FakeRunner(), FakeBuffer(),
                               update_policy=lambda batch_size, n_epochs: updates.append((batch_size, n_epochs)) or {"policy_loss": 0.5},
                               checkpoint=lambda lp, name: saved.append((name, lp.total_timesteps)), log=said.append, history=7)
    records = loop.train(total_timesteps=100, rollout_batch_size=2, rollout_batches=2, update_epochs=3, train_batch_size=16,
                         save_freq=60)
    # 4 episodes per iteration: lengths 10..14 cycle -> 46, 48, 50 ... timesteps per iteration; stop once >= 100
    per_iteration = [r["rollout"]["timesteps"] for r in records]
    assert per_iteration == [10 + 11 + 12 + 13, 14 + 10 + 11 + 12, 13 + 14 + 10 + 11]
    assert loop.total_timesteps == sum(per_iteration) >= 100 > sum(per_iteration[:-1])
    assert updates == [(16, 3)] * 3
    # checkpoint when >= save_freq timesteps since the last one (:712-714), final_model.pt at the end (:717)
    assert saved == [("checkpoint_2.pt", 93), ("final_model.pt", 141)]
    assert loop.last_save_timestep == 93 and records[1]["checkpoint"] == "checkpoint_2.pt" and "checkpoint" not in records[2]
    # statistics of THIS iteration's episodes; history bounded
    assert records[2]["rollout"]["mean_max_episode_reward"] == np.mean([8, 9, 10, 11])
    assert records[2]["rollout"]["max_episode_reward"] == 11.0
    assert records[2]["rollout"]["mean_episode_length"] == np.mean([13, 14, 10, 11])
    assert list(loop.episode_rewards) == [5.0, 6.0, 7.0, 8.0, 9.0, 10.0, 11.0]
    assert records[2]["mean_episode_reward_last_100"] == np.mean([5, 6, 7, 8, 9, 10, 11])
    assert said[0].startswith("Starting training from 0") and said[-1] == "Training completed!"


def test_resume_extends_or_completes(fake_collect):
    def make():
        return TL.PPOIterationLoop(FakeRunner(), FakeBuffer(), update_policy=lambda **k: {}, history=50)

    first = make()
    first.train(total_timesteps=90, rollout_batch_size=4, rollout_batches=1)
    state = first.state_dict()
    assert state["batch_runner"] == {"position": 8} and state["total_timesteps"] == 93

    extended = make()
    extended.load_state_dict(state)
    assert extended.resumed and extended.batch_runner.position == 8 and list(extended.episode_lengths) == list(first.episode_lengths)
    extended.train(total_timesteps=40, rollout_batch_size=4, rollout_batches=1, resume_extend_steps=True)
    assert extended.total_timesteps >= 93 + 40 and extended.batch_runner.position == 12

    absolute = make()
    absolute.load_state_dict(state)
    said = []
    absolute.log = said.append
    assert absolute.train(total_timesteps=90, rollout_batch_size=4, rollout_batches=1, resume_extend_steps=False) == []
    assert absolute.total_timesteps == 93 and any("No training needed" in s for s in said)

    # the continued run is the uninterrupted one
    straight = make()
    straight.train(total_timesteps=133, rollout_batch_size=4, rollout_batches=1)
    assert straight.total_timesteps == extended.total_timesteps
    assert list(straight.episode_rewards) == list(extended.episode_rewards)


def test_a_collection_without_timesteps_is_an_error(monkeypatch):
    monkeypatch.setattr(TL, "_collect", lambda *a, **k: {"episode_rewards": np.zeros(0, np.float32),
                                                         "episode_lengths": np.zeros(0, np.int64), "total_episodes": 0, "timesteps": 0})
    loop = TL.PPOIterationLoop(FakeRunner(), FakeBuffer(), update_policy=lambda **k: {})
    with pytest.raises(RuntimeError, match="no timesteps"):
        loop.train(total_timesteps=10)


def test_driving_a_trainer_shaped_object(fake_collect, monkeypatch):
    """for_trainer: the trainer's counters, history and writer see what its own collect_rollouts would have produced."""
    import torch

    from g2048.ppo import torch_action_wrapper as W

    class Writer:
        def __init__(self):
            self.scalars, self.closed = [], False

        def add_scalar(self, name, value, step):
            self.scalars.append((name, value, step))

        def close(self):
            self.closed = True

    class Agent(torch.nn.Module):
        pass

    class Trainer:
        def __init__(self):
            self.agent = Agent()
            self.batch_runner, self.rollout_buffer = FakeRunner(), FakeBuffer()
            self.gamma, self.lambda_gae, self.target_kl = 0.9, 0.8, 0.02
            self.max_samples_per_epoch, self.shuffle_on_reset = 20, True
            self.use_action_mask, self.device = True, torch.device("cpu")
            self.episode_rewards, self.episode_lengths = deque(maxlen=20), deque(maxlen=20)
            self.total_timesteps = self.total_epochs = self.total_update_steps = self.last_save_timestep = 0
            self.load_checkpoint_path = None
            self.writer = Writer()
            self.saved = []

        def update_policy(self, batch_size, n_epochs):
            self.agent.train()
            self.total_epochs += n_epochs
            self.total_update_steps += 5
            return {"policy_loss": 1.0, "n_updates": 5}

        def save_checkpoint(self, name):
            self.saved.append((name, self.total_timesteps, len(self.episode_rewards)))

    class FakeActionFunction:
        def __init__(self, agent, use_mask=False, device=None, autocast_dtype=None):
            self.agent, self.use_mask = agent, use_mask

    monkeypatch.setattr(W, "TorchActionFunction", FakeActionFunction)
    trainer = Trainer()
    loop = TL.PPOIterationLoop.for_trainer(trainer)
    act_fn = trainer.batch_runner.act_fn
    assert act_fn.agent is trainer.agent and act_fn.use_mask
    assert (loop.gamma, loop.lambda_gae, loop.target_kl, loop.max_samples_per_epoch) == (0.9, 0.8, 0.02, 20)
    trainer.agent.train()
    records = loop.train(total_timesteps=60, rollout_batch_size=3, rollout_batches=1, update_epochs=2, save_freq=30)
    assert trainer.batch_runner.act_fn is act_fn  # built once, not per collection
    assert len(records) == 2 and trainer.total_timesteps == loop.total_timesteps == 33 + 37
    assert trainer.total_epochs == loop.total_epochs == 4 and trainer.total_update_steps == loop.total_update_steps == 10
    assert list(trainer.episode_rewards) == list(loop.episode_rewards) == [0.0, 1.0, 2.0, 3.0, 4.0, 5.0]
    assert trainer.saved == [("checkpoint_1.pt", 33, 3), ("checkpoint_2.pt", 70, 6), ("final_model.pt", 70, 6)]
    assert trainer.writer.closed and ("rollout/mean_episode_length", 11.0, 33) in trainer.writer.scalars
    # a second loop over the same trainer continues from its counters
    again = TL.PPOIterationLoop.for_trainer(trainer)
    assert again.total_timesteps == 70 and len(again.episode_rewards) == 6
