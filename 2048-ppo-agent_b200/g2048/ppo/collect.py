"""The data path of ``PPOTrainer.collect_rollouts`` (src/ppo/ppo_trainer.py:154-249) without the trainer.

Per batch the reference runs ``run_actions_batch`` (numpy, reference format), one-hot encodes the actions in a
Python double loop (:197-202), stores the batch (:205-213) and walks every env once more for its episode
statistics (:218-227: ``episode_reward = max_t reward``, ``episode_length = first done + 1``).  Here the records stay
packed on the device: ``run_packed_batch`` -> ``RolloutBuffer.store_packed`` (the one-hot encoding only exists in
``get_buffer_data()`` if somebody asks for it), the two statistics are one reduction each.  With act_randomly /
act_drul the batch is ``run_flat_batch`` -> ``store_flat``: the recording play kernel writes the buffer's own layout.
"""
from __future__ import annotations

import numpy as np
import torch


def collect_rollouts(batch_runner, rollout_buffer, batch_size: int, num_batches: int, reset: bool = True) -> dict:
    """Plays ``num_batches`` batches of ``batch_size`` envs to termination with the runner's action function and
    appends the live steps to ``rollout_buffer``.

    Returns ``episode_rewards`` / ``episode_lengths`` (numpy, one entry per episode, in the reference's order: batch
    after batch, env after env), ``total_episodes`` and ``timesteps`` (= the buffer's size after the call).
    """
    if reset:
        rollout_buffer.reset()
    rewards, lengths = [], []
    with torch.no_grad():
        fused = getattr(batch_runner.act_fn, "policy_id", None) is not None
        for _ in range(num_batches):
            if fused:  # every env is played to termination by the lane that owns it: episode = its flat segment
                fr = batch_runner.run_flat_batch(batch_size)
                rollout_buffer.store_flat(fr)
                lengths.append(fr.lengths.long())
                rewards.append(fr.max_rewards)
                continue
            ro = batch_runner.run_packed_batch(batch_size)
            rollout_buffer.store_packed(ro)
            length = ro.lengths().long()
            length = torch.where(length == 0, torch.full_like(length, ro.t_steps), length)  # never terminated: num_steps
            rewards.append(ro.rewards[: ro.t_steps].max(dim=0).values)  # frozen steps carry reward 0 in the reference too
            lengths.append(length)
    episode_rewards = torch.cat(rewards).cpu().numpy() if rewards else np.zeros(0, np.float32)
    episode_lengths = torch.cat(lengths).cpu().numpy() if lengths else np.zeros(0, np.int64)
    return {"episode_rewards": episode_rewards, "episode_lengths": episode_lengths,
            "total_episodes": int(episode_lengths.shape[0]), "timesteps": int(rollout_buffer.buffer_size)}
