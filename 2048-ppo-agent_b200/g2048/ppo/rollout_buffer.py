"""RolloutBuffer -- the reference's src/ppo/rollout_buffer.py:4-206 backed by device memory.

The reference appends one Python object per step (store_batch, :164-187) and converts the lists
to arrays in get_buffer_data (:198-206; 2 017 bytes per step).  Here the buffer holds packed
records on the GPU -- bitboard 8 B + meta 1 B + reward/value/log-prob 12 B per step -- filled by
one compaction kernel per batch (``g2048_compact_records``: keep steps 0..first_done of every env,
env-major), and the reference-format arrays are produced on demand by ``g2048_expand_obs`` /
``g2048_unpack_flat_meta``.

``store_packed(rollout)`` takes a ``BatchRunner.run_packed_batch`` result without leaving the
device; ``store_batch(...)`` keeps the reference's numpy signature.

The reference class does not depend on the environment: any observation / action shape can be stored
(its own tests use 4- and 10-dimensional observations).  Batches that are not 2048 one-hot observations with
one-hot actions take the generic device path -- ``g2048_first_done_rows`` + ``g2048_compact_rows``, one contiguous
segment copy per env and field -- and are kept as float32 / bool rows; ``get_packed()`` is then unavailable.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import _native as N
from .. import engine as E


class RolloutBuffer:
    """Stores the trajectories of an environment/agent pair for PPO training."""

    def __init__(self, observation_dim: int, observation_length: int | list | tuple, action_dim: int) -> None:
        self.observation_dim = observation_dim
        self.observation_length = observation_length
        self.action_dim = action_dim
        self.reset()

    def reset(self):
        """Resets the buffer to its initial state."""
        # ("packed", (boards, meta, rewards, values, log_probs)) or ("generic", dict of the seven reference fields),
        # flat device tensors either way
        self._parts = []
        self.buffer_size = 0
        # the reference's per-step Python lists (rollout_buffer.py:47-56), kept as attributes for code that looks at
        # them; the data itself lives on the device and is read with get_buffer_data() / get_packed()
        self.observation_buffer = []
        self.action_buffer = []
        self.action_mask_buffer = []
        self.reward_buffer = []
        self.value_buffer = []
        self.log_prob_buffer = []
        self.termination_buffer = []

    # -- the reference's shape contract (rollout_buffer.py:58-126): the error TEXTS are part of its test suite -----
    def _obs_dims(self) -> tuple:
        if isinstance(self.observation_length, (tuple, list)):
            return (*self.observation_length, self.observation_dim)
        return (self.observation_length, self.observation_dim)

    def _validate_and_reshape_observations(self, observations: np.ndarray) -> np.ndarray:
        """(batch, time, ...) observations -> (batch, time, *observation dims); anything with the right number of
        elements per step is reshaped, anything else is refused with the reference's messages."""
        want = self._obs_dims()
        shape = observations.shape
        if len(shape) < 2:
            raise ValueError(
                f"Observations must have at least 2 dimensions (batch_size, time_steps, ...), "
                f"but got shape {shape}"
            )
        if shape[2:] == want:
            return observations
        if observations.size == 0:  # an empty batch: the reference trips over numpy's own reshape(-1) here
            try:
                observations.reshape(shape[0], shape[1], -1)
            except ValueError as e:
                raise ValueError(
                    f"Failed to reshape observations from shape {shape} to expected shape "
                    f"(batch_size, time_steps, {want}). Error: {str(e)}"
                )
        per_step, needed = int(np.prod(shape[2:], dtype=np.int64)), int(np.prod(want))
        if per_step != needed:
            # the reference raises this from inside a try block and wraps it once more: both texts, nested
            inner = (
                f"Cannot reshape observations from shape {shape} to expected shape "
                f"(batch_size, time_steps, {want}). "
                f"Flattened observations have {per_step} elements per timestep "
                f"but expected {needed} elements."
            )
            raise ValueError(
                f"Failed to reshape observations from shape {shape} to expected shape "
                f"(batch_size, time_steps, {want}). Error: {inner}"
            )
        return observations.reshape(shape[0], shape[1], *want)

    # -- storing --------------------------------------------------------------------------------
    def store_packed(self, rollout) -> int:
        """Append a ``PackedRollout`` (time-major device records).  Returns the steps kept."""
        t, b = rollout.t_steps, rollout.batch_size
        return self._compact(rollout.boards, rollout.meta, rollout.rewards, rollout.log_probs, rollout.values, t, b)

    def store_flat(self, rollout) -> int:
        """Append a ``FlatRollout`` (``BatchRunner.run_flat_batch``): the records are already laid out env after env,
        steps 0..first_done of each -- exactly what ``store_packed`` would have produced -- so nothing is copied."""
        total = int(rollout.boards.shape[0])
        if total == 0:
            return 0
        self._parts.append(("packed", (rollout.boards, rollout.meta, rollout.rewards, rollout.values, rollout.log_probs)))
        self.buffer_size += total
        return total

    def store_batch(self, observations, actions, action_masks, rewards, values, log_probs, terminations):
        """Reference signature (rollout_buffer.py:128-187): env-major (B, T, ...) numpy arrays.  2048 one-hot
        observations with one-hot (B,T,4) actions (or action indices (B,T)) are packed as bitboards; anything else
        is stored row by row through the generic compaction kernels."""
        observations = self._validate_and_reshape_observations(np.asarray(observations))
        dev = N.require_cuda()
        b, t = observations.shape[:2]
        if b == 0 or t == 0:
            return 0
        actions = np.asarray(actions)
        action_masks = np.asarray(action_masks)
        packable = (self.observation_dim == 31 and int(np.prod(observations.shape[2:-1])) == 16 and self.action_dim == 4
                    and action_masks.shape == (b, t, 4)
                    and (actions.shape == (b, t) and np.issubdtype(actions.dtype, np.integer)
                         or actions.shape == (b, t, 4) and bool((((actions == 0) | (actions == 1)).all(-1)
                                                                 & (actions.sum(-1) == 1)).all())))
        if packable:
            obs_t = torch.from_numpy(np.ascontiguousarray(observations)).to(dev)
            if obs_t.dtype not in (torch.bool, torch.uint8, torch.float32):
                obs_t = obs_t.to(torch.float32)
            boards = E.pack_obs(obs_t.contiguous())
            # bitboards hold one-hot cells only: anything else (test data, soft observations) stays as rows
            packable = bool(torch.equal(E.expand_obs(boards, torch.float32).view(obs_t.shape), obs_t.to(torch.float32)))
        if not packable:
            return self._store_rows(observations, actions, action_masks, rewards, values, log_probs, terminations, dev)
        boards = boards.view(b, t)
        act_idx = actions.argmax(-1) if actions.ndim == 3 else actions
        mask_bits = (action_masks.astype(np.uint8) * np.array([1, 2, 4, 8], np.uint8)).sum(-1)
        meta = (act_idx.astype(np.uint8) & 3) | (mask_bits.astype(np.uint8) << 2) | (np.asarray(terminations).astype(np.uint8) << 6)
        tm = lambda a, dt: torch.from_numpy(np.ascontiguousarray(np.asarray(a).T.astype(dt))).to(dev)  # noqa: E731
        return self._compact(
            boards.t().contiguous(), tm(meta, np.uint8), tm(rewards, np.float32),
            None if log_probs is None else tm(log_probs, np.float32),
            None if values is None else tm(values, np.float32), t, b,
        )

    def _store_rows(self, observations, actions, action_masks, rewards, values, log_probs, terminations, dev) -> int:
        """The generic path: every field is uploaded as it is (env-major) and the steps 0..first_done of each env are
        copied into flat device tensors (rollout_buffer.py:164-187), one segment per env and field."""
        b, t = observations.shape[:2]
        up = lambda a, dt: torch.from_numpy(np.ascontiguousarray(np.asarray(a).astype(dt, copy=False))).to(dev)  # noqa: E731
        zeros = np.zeros((b, t), np.float32)
        term = up(np.asarray(terminations).reshape(b, t) != 0, np.uint8)
        lengths = E.first_done_rows(term)
        offsets = E.exclusive_scan(lengths)
        total = int(offsets[-1].item())
        if total == 0:
            return 0
        fields = {
            "observations": up(observations, np.float32),
            "actions": up(actions, np.float32),
            "action_masks": up(action_masks != 0, np.uint8),
            "rewards": up(rewards, np.float32),
            "values": up(zeros if values is None else values, np.float32),
            "log_probs": up(zeros if log_probs is None else log_probs, np.float32),
            "terminations": term,
        }
        for name, x in fields.items():
            if x.shape[:2] != (b, t):
                raise ValueError(f"{name} must start with (batch_size, time_steps) = {(b, t)}, got {tuple(x.shape)}")
        self._parts.append(("generic", {k: E.compact_rows(x, lengths, offsets, total) for k, x in fields.items()}))
        self.buffer_size += total
        return total

    def _compact(self, rec_boards, rec_meta, rec_rewards, rec_log_probs, rec_values, t, b) -> int:
        dev = rec_meta.device
        lengths = E.episode_lengths(rec_meta, t, b)
        offsets = E.exclusive_scan(lengths)
        total = int(offsets[-1].item())
        if total == 0:
            return 0
        boards = torch.empty(total, dtype=torch.int64, device=dev)
        meta = torch.empty(total, dtype=torch.uint8, device=dev)
        rewards = torch.empty(total, dtype=torch.float32, device=dev)
        # a policy without log-probs / values (act_drul, act_randomly) stores zeros, like float32(None) would not
        log_probs = torch.zeros(total, dtype=torch.float32, device=dev)
        values = torch.zeros(total, dtype=torch.float32, device=dev)
        E.compact_records(rec_boards, rec_meta, rec_rewards, rec_log_probs, rec_values, t, b, lengths, offsets, 0,
                          boards, meta, rewards, log_probs, values)
        self._parts.append(("packed", (boards, meta, rewards, values, log_probs)))
        self.buffer_size += total
        return total

    # -- reading --------------------------------------------------------------------------------
    def get_packed(self) -> dict:
        """Flat device tensors: boards int64, meta uint8, rewards / values / log_probs float32."""
        if any(kind != "packed" for kind, _ in self._parts):
            raise ValueError("this buffer holds generic rows (not 2048 one-hot observations): there is no packed form, "
                             "use get_buffer_data()")
        if not self._parts:
            dev = N.require_cuda()
            z = lambda dt: torch.empty(0, dtype=dt, device=dev)  # noqa: E731
            return dict(boards=z(torch.int64), meta=z(torch.uint8), rewards=z(torch.float32),
                        values=z(torch.float32), log_probs=z(torch.float32))
        if len(self._parts) > 1:
            self._parts = [("packed", tuple(torch.cat([p[i] for _, p in self._parts]) for i in range(5)))]
        boards, meta, rewards, values, log_probs = self._parts[0][1]
        return dict(boards=boards, meta=meta, rewards=rewards, values=values, log_probs=log_probs)

    def _part_fields(self, kind: str, part) -> dict:
        """Device tensors of one part in the reference's layout."""
        if kind == "generic":
            return part
        boards, meta, rewards, values, log_probs = part
        onehot, masks, term = E.unpack_flat_meta(meta)
        return {"observations": E.expand_obs(boards, torch.float32).view(boards.shape[0], *self._obs_dims()),
                "actions": onehot, "action_masks": masks, "rewards": rewards, "values": values, "log_probs": log_probs,
                "terminations": term}

    def get_buffer_data(self):
        """Reference format (rollout_buffer.py:198-206): dict of numpy arrays."""
        if not self._parts:
            return {
                "observations": np.array([], dtype=np.float32), "actions": np.array([], dtype=np.float32),
                "action_masks": np.array([], dtype=bool), "rewards": np.array([], dtype=np.float32),
                "values": np.array([], dtype=np.float32), "log_probs": np.array([], dtype=np.float32),
                "terminations": np.array([], dtype=bool),
            }
        if all(kind == "packed" for kind, _ in self._parts):
            self.get_packed()  # one part
        parts = [self._part_fields(kind, part) for kind, part in self._parts]
        out = {}
        for name in ("observations", "actions", "action_masks", "rewards", "values", "log_probs", "terminations"):
            x = parts[0][name] if len(parts) == 1 else torch.cat([p[name] for p in parts])
            a = E.to_host(x)
            out[name] = a.astype(bool) if name in ("action_masks", "terminations") and a.dtype != bool else a
        return out
