"""GAE / returns / normalisation and the PPO dataset -- the reference's src/ppo/data_loader.py.

``compute_gae`` is the device path: flat packed buffer in, advantages and returns out, the
reverse recurrence of data_loader.py:103-130 run by ``g2048_gae_flat`` (bit-identical to the
reference's fp32 loop) and the global normalisation of :61-67 by ``g2048_normalize``.  A buffer
that is SHARDED over ranks is normalised with the global mean / std when the caller passes the
process group (``group=torch.distributed.group.WORLD``): the six fp64 moments are all-reduced.  The
collective is opt-in -- a buffer that lives on one rank of a multi-rank job (rank 0's evaluation
pass, a per-rank replay buffer) must not wait for peers that never call.

``PPODataset`` / ``create_ppo_dataloader`` keep the reference's constructor and item schema.
"""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch
from torch.utils.data import DataLoader, Dataset

from .. import _native as N
from .. import engine as E


def compute_gae(rewards: torch.Tensor, values: torch.Tensor, terminations: torch.Tensor, gamma: float = 0.99,
                lambda_gae: float = 0.95, normalize: bool = True, group=None, return_moments: bool = False):
    """Flat device buffers (N,) -> (advantages, returns) float32 device tensors.

    terminations: uint8 / bool, or the packed meta bytes (bit 6 = done) with ``terminations_is_meta``
    handled by the caller through ``meta_to_dones``.
    """
    dones = terminations if terminations.dtype == torch.uint8 else terminations.to(torch.uint8)
    adv, ret, moments = E.gae_flat(rewards.contiguous(), values.contiguous(), dones.contiguous(), gamma, lambda_gae)
    if (normalize or return_moments) and group is not None:
        from ..dist import allreduce_sum_

        allreduce_sum_(moments, group)
    if normalize:
        E.normalize_(adv, moments, 1)
        E.normalize_(ret, moments, 3)
    return (adv, ret, moments) if return_moments else (adv, ret)


def meta_to_dones(meta: torch.Tensor) -> torch.Tensor:
    """Packed meta bytes -> uint8 done flags."""
    return E.meta_dones(meta)


class DevicePPOBatches:
    """On-device minibatch pipeline over the packed buffer.

    Replaces ``PPODataset`` + ``DataLoader`` (data_loader.py:8-223) for the engine's own buffer format:
    GAE and normalisation run once on the device (``compute_gae``), every minibatch is produced by
    ``g2048_gather_minibatch`` -- the one-hot float observations exist only for the ``batch_size`` samples
    of the current step (the reference materialises 1 984 B per stored step up front).  The random subset
    per epoch (``max_samples_per_epoch`` / ``shuffle_on_reset``, :73-101) and the batch shuffle follow the
    reference's semantics; batches are dicts of DEVICE tensors with the item keys of ``PPODataset``
    (``actions`` are int64 indices -- the trainer's ``argmax`` of the one-hot, ppo_trainer.py:381).
    ``reuse_buffers=True`` writes the batches into two alternating sets of tensors (valid until the batch after the
    next).  ``obs_dtype=None`` leaves the observations out altogether: batches carry the 8-byte bitboards under
    ``boards`` for ``board_embedding.forward_from_boards`` (the embedding becomes a row gather).

    ``sample_records=True`` (default): the normalisation pass writes 32-byte sample records (``g2048_pack_samples``: board,
    meta, log-prob, value, normalised advantage and return of a step in ONE sector) and the minibatches are gathered
    from them -- one sector per sample instead of one per source array.  ``advantages`` / ``returns`` (the normalised
    flat arrays) are then produced on first use.
    ``epoch_prefetch=True``: ``__iter__`` gathers the WHOLE epoch (``len(self) * batch_size`` samples) with one launch
    into epoch-sized tensors (kept and reused by later epochs; ``max_samples_per_epoch`` x 2 012 B, 0.6 GB for the
    reference's 300 000) and yields views of them -- the per-minibatch cost drops from one launch (C4: ~14 us each from
    Python, 584 of them per iteration) to a slice.  A batch is valid until the next epoch starts.
    """

    EPOCH_PREFETCH_MAX_BYTES = 16 << 30

    def __init__(self, packed: dict, gamma: float = 0.99, lambda_gae: float = 0.95, batch_size: int = 32,
                 shuffle: bool = True, drop_last: bool = True, max_samples_per_epoch: int = None,
                 shuffle_on_reset: bool = False, obs_dtype=torch.float32, generator: torch.Generator = None,
                 group=None, reuse_buffers: bool = False, sample_records: bool = True, epoch_prefetch: bool = False):
        self.packed = packed
        self.batch_size = batch_size
        self.shuffle = shuffle
        self.drop_last = drop_last
        self.max_samples_per_epoch = max_samples_per_epoch
        self.shuffle_on_reset = shuffle_on_reset
        self.obs_dtype = obs_dtype
        self.generator = generator
        # reuse_buffers: full-size batches are written into two alternating sets of tensors instead of fresh ones (a
        # batch stays valid until the batch after the next is produced) -- saves eight allocations per minibatch
        self.reuse_buffers = reuse_buffers
        self._buffers, self._turn = [None, None], 0
        self.device = packed["rewards"].device
        self.total_length = packed["rewards"].shape[0]
        self.dones = meta_to_dones(packed["meta"]) if self.total_length else packed["meta"]
        self.epoch_prefetch = epoch_prefetch
        self._epoch_buffers = None
        self.records = None
        self._adv = self._ret = None
        if self.total_length and sample_records:
            # GAE (raw) -> moments (all-reduced over a sharded buffer) -> ONE pass that normalises and writes the records
            raw_adv, raw_ret, moments = compute_gae(packed["rewards"], packed["values"], self.dones, gamma, lambda_gae,
                                                    normalize=False, group=group, return_moments=True)
            self._raw = (raw_adv, raw_ret, moments)
            self.records = E.pack_samples(packed, raw_adv, raw_ret, moments)
        elif self.total_length:
            self._adv, self._ret = compute_gae(packed["rewards"], packed["values"], self.dones, gamma, lambda_gae,
                                               normalize=True, group=group)
        else:
            self._adv = self._ret = packed["rewards"]
            if group is not None:
                # an empty shard still takes part in the moment all-reduce of its peers (zeros), or they would wait forever
                from ..dist import allreduce_sum_

                allreduce_sum_(torch.zeros(6, dtype=torch.float64, device=self.device), group)
        if max_samples_per_epoch is None or max_samples_per_epoch >= self.total_length:
            self.length = self.total_length
            self.active_indices = None
        else:
            self.length = max_samples_per_epoch
            self.active_indices = self._sample_indices()

    def _normalised(self):
        if self._adv is None:  # records mode: the flat normalised arrays only exist if somebody asks for them
            raw_adv, raw_ret, moments = self._raw
            self._adv = E.normalize_(raw_adv.clone(), moments, 1)
            self._ret = E.normalize_(raw_ret.clone(), moments, 3)
        return self._adv, self._ret

    @property
    def advantages(self) -> torch.Tensor:
        return self._normalised()[0]

    @property
    def returns(self) -> torch.Tensor:
        return self._normalised()[1]

    def _randperm(self, n: int, m: int = None) -> torch.Tensor:
        """m (default n) distinct random positions of [0, n).  With an explicit torch generator: torch.randperm on
        the device.  Otherwise ``g2048_random_subset`` -- a keyed bijection evaluated at m points, O(m) instead of a
        sort of n keys (C4: 3.1e7 positions, 300 000 wanted: 2.8 ms -> a few microseconds) -- with the key drawn from
        torch's global CPU generator, so ``torch.manual_seed`` makes the epochs reproducible as in the reference."""
        m = n if m is None else m
        if self.generator is not None:
            return torch.randperm(n, device=self.device, generator=self.generator)[:m]
        key = torch.randint(0, 1 << 32, (2,), dtype=torch.int64)
        return E.random_subset(n, m, (int(key[0]), int(key[1])), self.device)

    def _sample_indices(self) -> torch.Tensor:
        return self._randperm(self.total_length, self.length)

    def reset_epoch(self):
        if self.shuffle_on_reset and self.active_indices is not None:
            self.active_indices = self._sample_indices()

    def __len__(self) -> int:
        if self.drop_last:
            return self.length // self.batch_size
        return (self.length + self.batch_size - 1) // self.batch_size

    def batch(self, indices: torch.Tensor) -> Dict[str, torch.Tensor]:
        """The minibatch of the given buffer positions (int64 device tensor)."""
        out = None
        if self.reuse_buffers and indices.shape[0] == self.batch_size:
            self._turn ^= 1
            if self._buffers[self._turn] is None:
                self._buffers[self._turn] = E.minibatch_buffers(self.batch_size, self.device, self.obs_dtype)
            out = self._buffers[self._turn]
        return self._gather(indices.contiguous(), out)

    def _gather(self, indices: torch.Tensor, out) -> Dict[str, torch.Tensor]:
        if self.records is not None:
            return E.gather_samples(indices, self.records, self.obs_dtype, out=out)
        return E.gather_minibatch(indices, self.packed, self.advantages, self.returns, self.obs_dtype, out=out)

    def _bytes_per_sample(self) -> int:
        obs = 0 if self.obs_dtype is None else 496 * torch.empty((), dtype=self.obs_dtype).element_size()
        return obs + 8 + 4 + 16 + (8 if self.obs_dtype is None else 0)

    def __iter__(self):
        order = self._randperm(self.length) if self.shuffle else torch.arange(self.length, device=self.device)
        if self.active_indices is not None:
            order = self.active_indices[order]
        n_batches = len(self)
        used = min(self.length, n_batches * self.batch_size)
        if self.epoch_prefetch and used and used * self._bytes_per_sample() <= self.EPOCH_PREFETCH_MAX_BYTES:
            if self._epoch_buffers is None:
                self._epoch_buffers = E.minibatch_buffers(used, self.device, self.obs_dtype)
            epoch = self._gather(order[:used].contiguous(), self._epoch_buffers)
            keys = list(epoch)
            parts = [epoch[k].split(self.batch_size) for k in keys]  # views, one C++ call per field
            for views in zip(*parts):
                yield dict(zip(keys, views))
            return
        for b in range(n_batches):
            yield self.batch(order[b * self.batch_size: (b + 1) * self.batch_size])


class _LazySamples:
    """What ``PPODataset.__getitems__`` hands to the DataLoader's collate function: the positions of one batch.
    ``_collate`` (the collate function of ``create_ppo_dataloader``) turns it into the batch with one indexed read
    per field; to any other collate function it looks like the list of per-sample dicts ``__getitem__`` returns."""

    def __init__(self, dataset: "PPODataset", positions: torch.Tensor):
        self.dataset = dataset
        self.positions = positions

    def __len__(self) -> int:
        return self.positions.shape[0]

    def __getitem__(self, i):
        return self.dataset._item(self.positions[i])

    def __iter__(self):
        return (self.dataset._item(p) for p in self.positions)

    def collated(self) -> Dict[str, torch.Tensor]:
        return self.dataset._item(self.positions)


def _collate(batch):
    """Collate function of ``create_ppo_dataloader``: same batches as torch's default collation of the per-sample
    dicts (data_loader.py:217-223 builds a plain DataLoader), without the per-sample Python work -- a 2 048-sample
    batch is nine indexed reads instead of 2 048 dicts of nine tensors each."""
    if isinstance(batch, _LazySamples):
        return batch.collated()
    return torch.utils.data.default_collate(batch)


class PPODataset(Dataset):
    """Dataset over RolloutBuffer data with GAE advantages and returns (data_loader.py:8-166)."""

    def __init__(self, buffer_data: Dict[str, np.ndarray], gamma: float = 0.99, lambda_gae: float = 0.95,
                 max_samples_per_epoch: int = None, shuffle_on_reset: bool = False):
        self.gamma = gamma
        self.lambda_gae = lambda_gae
        self.max_samples_per_epoch = max_samples_per_epoch
        self.shuffle_on_reset = shuffle_on_reset

        self.observations = torch.from_numpy(buffer_data["observations"]).float()
        self.actions = torch.from_numpy(buffer_data["actions"]).float()
        self.action_masks = torch.from_numpy(buffer_data["action_masks"]).bool()
        self.rewards = torch.from_numpy(buffer_data["rewards"]).float()
        self.values = torch.from_numpy(buffer_data["values"]).float()
        self.log_probs = torch.from_numpy(buffer_data["log_probs"]).float()
        self.terminations = torch.from_numpy(buffer_data["terminations"]).bool()

        # advantages and returns: ONE device pass for the recurrence (bit-identical to the reference's loop), then the
        # reference's own host-side normalisation -- the same torch float32 mean / std expressions on the same values
        # (data_loader.py:61-67), so the normalised arrays are equal to the reference's as well, not merely within 1e-5
        self.raw_advantages, self.raw_returns = self._compute_gae_returns()
        if len(self.rewards):
            self.advantages = (self.raw_advantages - self.raw_advantages.mean()) / (self.raw_advantages.std() + 1e-8)
            self.returns = (self.raw_returns - self.raw_returns.mean()) / (self.raw_returns.std() + 1e-8)
        else:
            self.advantages, self.returns = self.raw_advantages, self.raw_returns

        self.total_length = len(self.observations)
        if self.max_samples_per_epoch is None or self.max_samples_per_epoch >= self.total_length:
            self.length = self.total_length
            self.active_indices = None
        else:
            self.length = self.max_samples_per_epoch
            self.active_indices = self._sample_indices()

    def _sample_indices(self) -> torch.Tensor:
        return torch.randperm(self.total_length)[: self.length]

    def reset_epoch(self):
        if self.shuffle_on_reset and self.active_indices is not None:
            self.active_indices = self._sample_indices()

    def _compute_gae_returns(self):
        """Un-normalised advantages / returns (data_loader.py:103-130) through the CUDA kernel."""
        if len(self.rewards) == 0:
            return torch.zeros_like(self.rewards), torch.zeros_like(self.rewards)
        adv, ret = E.gae_host(self.rewards.numpy(), self.values.numpy(), self.terminations.numpy().astype(np.uint8),
                              self.gamma, self.lambda_gae, False)
        return torch.from_numpy(adv), torch.from_numpy(ret)

    def __len__(self) -> int:
        return self.length

    def __getitems__(self, indices) -> _LazySamples:
        """Batched fetch (torch's DataLoader calls this with the indices of one batch)."""
        idx = torch.as_tensor(indices, dtype=torch.long)
        return _LazySamples(self, self.active_indices[idx] if self.active_indices is not None else idx)

    def __getitem__(self, idx: int) -> Dict[str, torch.Tensor]:
        actual_idx = self.active_indices[idx] if self.active_indices is not None else idx
        return self._item(actual_idx)

    def _item(self, actual_idx) -> Dict[str, torch.Tensor]:
        """The fields at buffer position(s) actual_idx (an int or an index tensor)."""
        return {
            "observations": self.observations[actual_idx],
            "actions": self.actions[actual_idx],
            "action_masks": self.action_masks[actual_idx],
            "rewards": self.rewards[actual_idx],
            "values": self.values[actual_idx],
            "log_probs": self.log_probs[actual_idx],
            "terminations": self.terminations[actual_idx],
            "advantages": self.advantages[actual_idx],
            "returns": self.returns[actual_idx],
        }


def create_ppo_dataloader(buffer_data: Dict[str, np.ndarray], gamma: float = 0.99, lambda_gae: float = 0.95,
                          batch_size: int = 32, shuffle: bool = True, drop_last: bool = True, num_workers: int = 0,
                          max_samples_per_epoch: int = None, shuffle_on_reset: bool = False) -> DataLoader:
    """Same factory as data_loader.py:169-223."""
    dataset = PPODataset(buffer_data, gamma=gamma, lambda_gae=lambda_gae,
                         max_samples_per_epoch=max_samples_per_epoch, shuffle_on_reset=shuffle_on_reset)
    return DataLoader(dataset, batch_size=batch_size, shuffle=shuffle, drop_last=drop_last, num_workers=num_workers,
                      collate_fn=_collate)
