"""The orchestration of ``PPOTrainer.train`` (src/ppo/ppo_trainer.py:625-728) around a learner it does not own.

One iteration of the reference is ``collect_rollouts`` (:154-249) then ``update_policy`` (:316-500), repeated until
``total_timesteps`` reaches its target, with a checkpoint every ``save_freq`` timesteps.  Everything in it except
the network's forward / backward / optimizer step is data movement, and that part is what this class runs:

* the rollouts go through ``collect.collect_rollouts`` (recording play kernel or the fused policy step, straight
  into the buffer's packed layout);
* the episode history is a bounded deque as in the reference (:140-146), but its statistics are taken from a list
  of its tail -- the reference slices the deque itself (:237-239, :706) and raises ``TypeError`` on the first
  iteration;
* the update walks ``DevicePPOBatches`` (GAE, normalisation, 32-byte sample records, one gather per epoch) and
  hands every minibatch to ``minibatch_step`` -- the caller's loss + backward + optimizer step.  The per-minibatch
  metrics stay device tensors until the epoch ends: one host read per epoch (the KL early stop, :452-458) instead of
  the reference's five ``.item()`` per minibatch (:434-446);
* ``state_dict`` carries the counters, the history AND the runner's key chain, so a resumed run plays the same
  episodes as an uninterrupted one (the reference's checkpoint, :511-530, restarts the chain from its seed).

An existing trainer object (the reference's ``PPOTrainer`` with its own ``update_policy``) can be driven instead of a
``minibatch_step``: ``PPOIterationLoop.for_trainer(trainer)``; its counters and history are kept in step.
"""
from __future__ import annotations

from collections import deque
from typing import Callable, Dict, List, Optional

import numpy as np
import torch

from .collect import collect_rollouts as _collect

METRIC_KEYS = ("policy_loss", "value_loss", "entropy_loss", "total_loss")


def tail(history, count: int) -> list:
    """The last ``count`` entries of a deque (or any sequence) as a list: ``deque[-k:]`` is a TypeError."""
    count = max(0, min(int(count), len(history)))
    if count == 0:
        return []
    items = list(history)
    return items[len(items) - count:]


class PPOIterationLoop:
    """collect -> update -> checkpoint, until ``total_timesteps`` is reached.

    ``minibatch_step(batch) -> dict`` receives the dict ``DevicePPOBatches`` yields (device tensors: observations or
    boards, actions as int64 INDICES -- what the trainer's argmax at :368 produces --, action_masks, log_probs, values,
    normalised advantages and returns) and returns any of ``policy_loss``,
    ``value_loss``, ``entropy_loss``, ``total_loss``, ``kl`` as floats or 0-d tensors (``kl`` = mean of
    old_log_prob - new_log_prob, :443).  ``update_policy(batch_size=, n_epochs=) -> dict`` replaces the whole update
    (the reference trainer's bound method).  ``checkpoint(loop, name)`` is called where the reference saves.  With a
    runner that owns a shard of every batch (``BatchRunner(shard=(rank, world))``) pass the ranks' process ``group``: the
    buffer's advantages and returns are then normalised with the moments of ALL shards (one all-reduce of six numbers).
    """

    def __init__(self, batch_runner, rollout_buffer, minibatch_step: Optional[Callable] = None,
                 update_policy: Optional[Callable] = None, gamma: float = 0.99, lambda_gae: float = 0.95,
                 target_kl: float = 0.01, max_samples_per_epoch: Optional[int] = None, shuffle_on_reset: bool = False,
                 obs_dtype=torch.float32, checkpoint: Optional[Callable] = None, log: Optional[Callable] = None,
                 history: Optional[int] = None, agent=None, group=None):
        if (minibatch_step is None) == (update_policy is None):
            raise ValueError("give exactly one of minibatch_step and update_policy")
        self.batch_runner = batch_runner
        self.rollout_buffer = rollout_buffer
        self.minibatch_step = minibatch_step
        self._update_policy = update_policy
        self.gamma, self.lambda_gae, self.target_kl = gamma, lambda_gae, target_kl
        self.max_samples_per_epoch, self.shuffle_on_reset = max_samples_per_epoch, shuffle_on_reset
        self.obs_dtype = obs_dtype
        self.checkpoint, self.log = checkpoint, log
        self.trainer = None
        self.agent = agent  # the policy network, if the rollouts use one: put in eval mode before every collection
        self.group = group  # process group of a buffer SHARDED over ranks: advantages / returns are normalised with the global moments
        if history is None:  # :140-143
            history = max_samples_per_epoch if max_samples_per_epoch is not None else 10000
        self.episode_rewards: deque = deque(maxlen=history)
        self.episode_lengths: deque = deque(maxlen=history)
        self.total_timesteps = 0
        self.total_epochs = 0
        self.total_update_steps = 0
        self.last_save_timestep = 0
        self.resumed = False

    @classmethod
    def for_trainer(cls, trainer, autocast_dtype=None, **kwargs) -> "PPOIterationLoop":
        """Drives an object shaped like the reference's PPOTrainer: its runner, buffer, agent, hyper-parameters and
        ``update_policy``; its own ``collect_rollouts`` / ``train`` are not called.  The runner's action function
        is built ONCE around the trainer's agent (the reference wraps -- and re-traces -- the agent at the top of every
        collection, :170-176; the module is updated in place, so one wrapper sees every new set of weights)."""
        from .torch_action_wrapper import TorchActionFunction

        act_fn = getattr(trainer.batch_runner, "act_fn", None)
        if not (isinstance(act_fn, TorchActionFunction) and act_fn.agent is trainer.agent):
            trainer.batch_runner.act_fn = TorchActionFunction(trainer.agent, use_mask=trainer.use_action_mask,
                                                              device=trainer.device, autocast_dtype=autocast_dtype)
        kwargs.setdefault("agent", trainer.agent)
        loop = cls(trainer.batch_runner, trainer.rollout_buffer, update_policy=trainer.update_policy,
                   gamma=trainer.gamma, lambda_gae=trainer.lambda_gae, target_kl=trainer.target_kl,
                   max_samples_per_epoch=trainer.max_samples_per_epoch, shuffle_on_reset=trainer.shuffle_on_reset,
                   history=trainer.episode_rewards.maxlen, **kwargs)
        loop.trainer = trainer
        loop.total_timesteps = int(trainer.total_timesteps)
        loop.total_epochs = int(trainer.total_epochs)
        loop.total_update_steps = int(trainer.total_update_steps)
        loop.last_save_timestep = int(trainer.last_save_timestep)
        loop.episode_rewards.extend(trainer.episode_rewards)
        loop.episode_lengths.extend(trainer.episode_lengths)
        loop.resumed = getattr(trainer, "load_checkpoint_path", None) is not None
        if loop.checkpoint is None and hasattr(trainer, "save_checkpoint"):
            loop.checkpoint = lambda _loop, name: trainer.save_checkpoint(name)
        return loop

    def _say(self, message: str) -> None:
        if self.log is not None:
            self.log(message)

    def _mirror(self) -> None:
        t = self.trainer
        if t is None:
            return
        t.total_timesteps, t.last_save_timestep = self.total_timesteps, self.last_save_timestep
        if self._update_policy is None:  # otherwise the trainer's own update_policy counts these itself
            t.total_epochs, t.total_update_steps = self.total_epochs, self.total_update_steps

    # ---- rollouts (:154-249) -----------------------------------------------------------------------------------------
    def collect_rollouts(self, batch_size: int, num_batches: int) -> Dict[str, float]:
        """Refills the buffer with ``num_batches`` batches of ``batch_size`` episodes and returns the three rollout
        scalars the reference logs (:236-249), computed over THIS call's episodes."""
        if self.agent is not None:
            self.agent.eval()  # :167-168; update_policy leaves the agent in training mode (:345)
        out = _collect(self.batch_runner, self.rollout_buffer, batch_size, num_batches)
        rewards, lengths = out["episode_rewards"], out["episode_lengths"]
        self.episode_rewards.extend(float(r) for r in rewards)
        self.episode_lengths.extend(int(n) for n in lengths)
        if self.trainer is not None:
            self.trainer.episode_rewards.extend(float(r) for r in rewards)
            self.trainer.episode_lengths.extend(int(n) for n in lengths)
        self.total_timesteps += out["timesteps"]
        self._mirror()
        stats = {"timesteps": out["timesteps"], "total_episodes": out["total_episodes"]}
        recent_r, recent_n = tail(self.episode_rewards, out["total_episodes"]), tail(self.episode_lengths, out["total_episodes"])
        if recent_r:
            stats["mean_max_episode_reward"] = float(np.mean(recent_r))
            stats["max_episode_reward"] = float(np.max(recent_r))
            stats["mean_episode_length"] = float(np.mean(recent_n))
            writer = getattr(self.trainer, "writer", None)
            if writer is not None:
                for name in ("mean_max_episode_reward", "max_episode_reward", "mean_episode_length"):
                    writer.add_scalar(f"rollout/{name}", stats[name], self.total_timesteps)
        self._say(f"Collected {out['timesteps']} timesteps from {out['total_episodes']} episodes")
        return stats

    # ---- update (:316-500) -------------------------------------------------------------------------------------------
    def update_policy(self, batch_size: int = 64, n_epochs: int = 4) -> Dict[str, float]:
        if self._update_policy is not None:
            metrics = self._update_policy(batch_size=batch_size, n_epochs=n_epochs)
            if self.trainer is not None:
                self.total_epochs, self.total_update_steps = int(self.trainer.total_epochs), int(self.trainer.total_update_steps)
            return metrics
        if self.rollout_buffer.buffer_size == 0:
            self._say("No data in rollout buffer")
            return {}
        from .data_loader import DevicePPOBatches

        batches = DevicePPOBatches(self.rollout_buffer.get_packed(), gamma=self.gamma, lambda_gae=self.lambda_gae,
                                   batch_size=batch_size, shuffle=True, drop_last=True,
                                   max_samples_per_epoch=self.max_samples_per_epoch, shuffle_on_reset=self.shuffle_on_reset,
                                   obs_dtype=self.obs_dtype, sample_records=True, epoch_prefetch=True, group=self.group)
        device = batches.device
        sums = torch.zeros(len(METRIC_KEYS), dtype=torch.float64, device=device)
        n_updates, mean_kl = 0, 0.0
        for epoch in range(n_epochs):
            batches.reset_epoch()
            kl_sum = torch.zeros((), dtype=torch.float64, device=device)
            epoch_batches = 0
            for batch in batches:
                out = self.minibatch_step(batch) or {}
                for i, key in enumerate(METRIC_KEYS):
                    if key in out:
                        sums[i] += _scalar(out[key], device)
                if "kl" in out:
                    kl_sum += _scalar(out["kl"], device)
                epoch_batches += 1
            n_updates += epoch_batches
            self.total_update_steps += epoch_batches
            self.total_epochs += 1
            mean_kl = float(kl_sum) / epoch_batches if epoch_batches else 0.0  # the epoch's one host read
            if mean_kl > self.target_kl:
                self._say(f"Early stopping at epoch {epoch} due to high KL divergence: {mean_kl:.6f}")
                break
        host = sums.cpu().tolist()
        metrics = {key: (host[i] / n_updates if n_updates else 0) for i, key in enumerate(METRIC_KEYS)}
        metrics["kl_divergence"] = mean_kl
        metrics["n_updates"] = n_updates
        self._mirror()
        return metrics

    # ---- the loop (:625-728) -----------------------------------------------------------------------------------------
    def train(self, total_timesteps: int, rollout_batch_size: int = 32, rollout_batches: int = 4, update_epochs: int = 4,
              train_batch_size: int = 64, save_freq: int = 10000, resume_extend_steps: bool = True) -> List[dict]:
        """Same arguments and stopping rule as the reference; returns one record per iteration (rollout scalars,
        update metrics, ``mean_episode_reward_last_100``) instead of None."""
        start = self.total_timesteps
        target = start + total_timesteps if resume_extend_steps else total_timesteps
        self._say(f"{'Resuming' if self.resumed else 'Starting'} training from {start} timesteps to reach {target}")
        records: List[dict] = []
        if start >= target:
            self._say(f"Already trained for {start} timesteps, target is {target}. No training needed.")
            return records
        iteration = 0
        while self.total_timesteps < target:
            iteration += 1
            before = self.total_timesteps
            record = {"iteration": iteration, "rollout": self.collect_rollouts(rollout_batch_size, rollout_batches)}
            if self.total_timesteps == before:
                raise RuntimeError("collect_rollouts added no timesteps: the loop would never reach its target")
            record["update"] = self.update_policy(batch_size=train_batch_size, n_epochs=update_epochs)
            record["timesteps"] = self.total_timesteps
            recent = tail(self.episode_rewards, 100)
            if recent:
                record["mean_episode_reward_last_100"] = float(np.mean(recent))
            self._say(f"Iteration {iteration}, Timesteps: {self.total_timesteps}/{target}")
            if self.total_timesteps - self.last_save_timestep >= save_freq:
                self.last_save_timestep = self.total_timesteps
                self._mirror()
                if self.checkpoint is not None:
                    self.checkpoint(self, f"checkpoint_{iteration}.pt")
                record["checkpoint"] = f"checkpoint_{iteration}.pt"
            records.append(record)
        self._say("Training completed!")
        if self.checkpoint is not None:
            self.checkpoint(self, "final_model.pt")
        writer = getattr(self.trainer, "writer", None)
        if writer is not None:
            writer.close()
        return records

    # ---- resume ------------------------------------------------------------------------------------------------------
    def state_dict(self) -> dict:
        state = {"total_timesteps": self.total_timesteps, "total_epochs": self.total_epochs,
                 "total_update_steps": self.total_update_steps, "last_save_timestep": self.last_save_timestep,
                 "episode_rewards": list(self.episode_rewards), "episode_lengths": list(self.episode_lengths)}
        if hasattr(self.batch_runner, "state_dict"):
            state["batch_runner"] = self.batch_runner.state_dict()
        return state

    def load_state_dict(self, state: dict) -> None:
        self.total_timesteps = int(state.get("total_timesteps", 0))
        self.total_epochs = int(state.get("total_epochs", 0))
        self.total_update_steps = int(state.get("total_update_steps", 0))
        self.last_save_timestep = int(state.get("last_save_timestep", 0))
        self.episode_rewards = deque(state.get("episode_rewards", []), maxlen=self.episode_rewards.maxlen)
        self.episode_lengths = deque(state.get("episode_lengths", []), maxlen=self.episode_lengths.maxlen)
        if "batch_runner" in state and hasattr(self.batch_runner, "load_state_dict"):
            self.batch_runner.load_state_dict(state["batch_runner"])
        self.resumed = True
        self._mirror()


def _scalar(value, device) -> torch.Tensor:
    if isinstance(value, torch.Tensor):
        return value.detach().to(device=device, dtype=torch.float64).reshape(())
    return torch.tensor(float(value), dtype=torch.float64, device=device)
