"""PPOIterationLoop on the GPU: whole iterations (recorded rollouts -> GAE -> sample records -> minibatches -> a
learner's step) with the built-in random policy and with a small policy network, the per-epoch KL stop, resume."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


class TinyAgent(torch.nn.Module):
    """forward(obs (B,16,31), mask) -> (logits (B,4), values (B,1)): the PPOAgent contract (src/ppo/ppo_agent.py:90-125)."""

    def __init__(self):
        super().__init__()
        self.body = torch.nn.Linear(16 * 31, 32)
        self.policy = torch.nn.Linear(32, 4)
        self.value = torch.nn.Linear(32, 1)

    def forward(self, obs, mask=None):
        h = torch.tanh(self.body(obs.reshape(obs.shape[0], -1).float()))
        return self.policy(h), self.value(h)


def ppo_step(agent, optimizer, clip=0.2):
    """The arithmetic of PPOTrainer._compute_ppo_loss + the optimizer step (src/ppo/ppo_trainer.py:251-314,389-430)."""
    def step(batch):
        agent.train()  # :345
        logits, values = agent(batch["observations"])
        logits = logits - 1e8 * (1.0 - batch["action_masks"].float())
        log_p = torch.log_softmax(logits, dim=-1)
        new_lp = log_p.gather(1, batch["actions"].unsqueeze(1)).squeeze(1)  # action INDICES (the trainer argmaxes, :368)
        ratio = torch.exp(new_lp - batch["log_probs"])
        adv = batch["advantages"]
        policy_loss = -torch.min(ratio * adv, torch.clamp(ratio, 1 - clip, 1 + clip) * adv)
        value_loss = (values.flatten() - batch["returns"]) ** 2
        entropy_loss = (log_p.exp() * log_p.clamp_min(-50)).sum(-1)
        loss = (policy_loss + 0.5 * value_loss + 0.01 * entropy_loss).mean()
        optimizer.zero_grad()
        loss.backward()
        optimizer.step()
        return {"policy_loss": policy_loss.mean(), "value_loss": value_loss.mean(), "entropy_loss": entropy_loss.mean(),
                "total_loss": loss, "kl": (batch["log_probs"] - new_lp.detach()).mean()}
    return step


def test_iterations_with_the_random_policy_count_like_the_reference():
    import g2048
    from g2048.ppo.train_loop import PPOIterationLoop

    seen = []

    def step(batch):
        seen.append({k: v.shape for k, v in batch.items()})
        assert batch["observations"].is_cuda and batch["observations"].dtype == torch.float32
        one_hot = batch["observations"].sum(-1)
        assert bool((one_hot == 1).all())  # every cell of every sampled board is one-hot
        return {"policy_loss": 2.0, "value_loss": torch.tensor(4.0, device="cuda"), "kl": 0.0}

    runner = g2048.BatchRunner(init_seed=3, act_fn=g2048.act_randomly)
    buffer = g2048.RolloutBuffer(31, 16, 4)
    saved = []
    loop = PPOIterationLoop(runner, buffer, minibatch_step=step, max_samples_per_epoch=5000, shuffle_on_reset=True,
                            checkpoint=lambda lp, name: saved.append(name))
    records = loop.train(total_timesteps=100000, rollout_batch_size=256, rollout_batches=2, update_epochs=3,
                         train_batch_size=512, save_freq=10 ** 9)
    assert loop.total_timesteps >= 100000 and records[-1]["timesteps"] == loop.total_timesteps
    assert sum(r["rollout"]["timesteps"] for r in records) == loop.total_timesteps
    per_epoch = 5000 // 512  # drop_last
    assert len(seen) == len(records) * 3 * per_epoch == loop.total_update_steps and loop.total_epochs == 3 * len(records)
    assert seen[0]["observations"] == (512, 16, 31) and seen[0]["advantages"] == (512,)
    for r in records:
        ro, up = r["rollout"], r["update"]
        assert ro["total_episodes"] == 512 and ro["mean_episode_length"] == pytest.approx(ro["timesteps"] / 512)
        assert 60 < ro["mean_episode_length"] < 200 and ro["max_episode_reward"] >= ro["mean_max_episode_reward"] > 16
        assert up == {"policy_loss": 2.0, "value_loss": 4.0, "entropy_loss": 0.0, "total_loss": 0.0, "kl_divergence": 0.0,
                      "n_updates": 3 * per_epoch}
    assert saved == ["final_model.pt"] and len(loop.episode_rewards) == min(5000, 512 * len(records))
    # the statistics are those of the episodes in the buffer: lengths sum to its size, rewards are the per-episode maxima
    packed = buffer.get_packed()
    assert int(sum(g2048.ppo.train_loop.tail(loop.episode_lengths, 512))) == packed["rewards"].shape[0] == buffer.buffer_size
    assert max(g2048.ppo.train_loop.tail(loop.episode_rewards, 512)) == float(packed["rewards"].max())


def test_kl_stops_the_epochs_early():
    import g2048
    from g2048.ppo.train_loop import PPOIterationLoop

    calls = []
    runner = g2048.BatchRunner(init_seed=1, act_fn=g2048.act_drul)
    loop = PPOIterationLoop(runner, g2048.RolloutBuffer(31, 16, 4), target_kl=0.05,
                            minibatch_step=lambda b: calls.append(1) or {"kl": torch.tensor(0.02 * (1 + len(calls) // 10), device="cuda")})
    loop.collect_rollouts(64, 1)
    n_batches = loop.rollout_buffer.buffer_size // 1024
    assert n_batches >= 5
    metrics = loop.update_policy(batch_size=1024, n_epochs=50)
    # kl per minibatch = 0.02 * (1 + calls // 10): the first epoch whose MEAN exceeds 0.05 ends the update (:452-458)
    kls = [0.02 * (1 + (i + 1) // 10) for i in range(len(calls))]
    epochs = len(calls) // n_batches
    assert len(calls) == epochs * n_batches and epochs < 50 and loop.total_epochs == epochs
    means = [np.mean(kls[e * n_batches:(e + 1) * n_batches]) for e in range(epochs)]
    assert all(m <= 0.05 for m in means[:-1]) and means[-1] > 0.05
    assert metrics["kl_divergence"] == pytest.approx(means[-1]) and metrics["n_updates"] == len(calls)


def test_a_network_policy_learns_through_the_loop_and_resumes_bit_identically():
    import g2048
    from g2048.ppo.train_loop import PPOIterationLoop

    def make(seed_agent=0):
        torch.manual_seed(seed_agent)
        agent = TinyAgent().cuda()
        optimizer = torch.optim.Adam(agent.parameters(), lr=1e-3)
        act = g2048.TorchActionFunction(agent, use_mask=True, device=torch.device("cuda"))
        runner = g2048.BatchRunner(init_seed=11, act_fn=act)
        loop = PPOIterationLoop(runner, g2048.RolloutBuffer(31, 16, 4), minibatch_step=ppo_step(agent, optimizer), agent=agent,
                                target_kl=1e9)
        return agent, optimizer, loop

    def iteration(loop):
        torch.manual_seed(100 + loop.total_epochs)  # the minibatch order of this iteration
        ro = loop.collect_rollouts(128, 2)
        assert not loop.agent.training  # :167-168
        up = loop.update_policy(batch_size=1024, n_epochs=2)
        return ro, up, {k: v.clone() for k, v in loop.rollout_buffer.get_packed().items()}

    agent, optimizer, loop = make()
    before = [p.detach().clone() for p in agent.parameters()]
    first = iteration(loop)
    assert agent.training  # the update leaves it so; the next collection puts it back in eval mode
    checkpoint = {"loop": loop.state_dict(), "agent": {k: v.clone() for k, v in agent.state_dict().items()},
                  "optimizer": optimizer.state_dict()}
    import copy

    checkpoint = copy.deepcopy(checkpoint)
    second = iteration(loop)
    assert any(not torch.equal(a, b.detach()) for a, b in zip(before, agent.parameters()))
    for ro, up, _ in (first, second):
        assert ro["total_episodes"] == 256 and up["n_updates"] == 2 * (ro["timesteps"] // 1024) > 0
        assert all(np.isfinite(v) for v in up.values())
    # the second collection used the UPDATED weights: its records differ from a replay of the first
    assert first[2]["boards"].shape != second[2]["boards"].shape or not torch.equal(first[2]["boards"], second[2]["boards"])

    # resume from the checkpoint in fresh objects: same second iteration, bit for bit (key chain included)
    agent2, optimizer2, loop2 = make(seed_agent=5)
    agent2.load_state_dict(checkpoint["agent"])
    optimizer2.load_state_dict(checkpoint["optimizer"])
    loop2.load_state_dict(checkpoint["loop"])
    assert loop2.total_timesteps == first[0]["timesteps"] and loop2.resumed
    resumed = iteration(loop2)
    assert resumed[0] == second[0]
    for k in second[2]:
        assert torch.equal(second[2][k], resumed[2][k]), k
    assert resumed[1] == second[1]
    for a, b in zip(agent.parameters(), agent2.parameters()):
        assert torch.equal(a, b)
