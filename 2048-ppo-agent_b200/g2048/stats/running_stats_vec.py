"""Running mean / variance per feature row -- the reference's src/stats/running_stats_vec.py:4-99.

``push`` reduces the new samples on the device (``g2048_row_moments``: count, mean, population
variance per row in fp64) and folds the triple into the running one with the same Chan update
the reference uses (:74-87).  ``merge`` / ``all_reduce`` fold in other triples, which is how
episode statistics are combined across GPUs.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import _native as N
from .. import engine as E


class RunningStatsVec(object):
    """Running mean and variance computation for vectors."""

    def __init__(self):
        self.clear()

    def clear(self):
        """Reset the number of samples, mean and variance to zero."""
        self.num_samples = np.zeros((1, 1), dtype=np.int64)
        self._mean = np.zeros((1, 1), dtype=np.float64)
        self._variance = np.zeros((1, 1), dtype=np.float64)

    def push(self, x):
        """x: (num_features, num_samples) numpy array or torch tensor."""
        if len(x.shape) != 2:
            raise ValueError("Input array should have 2 dimensions.")
        dev = N.require_cuda()
        if isinstance(x, torch.Tensor):
            t = x.to(device=dev, dtype=torch.float64).contiguous()
        else:
            t = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64)).to(dev)
        trip = E.row_moments(t).cpu().numpy()
        self.merge_triple(trip[:, 0:1].astype(np.int64), trip[:, 1:2], trip[:, 2:3])

    def merge_triple(self, n_b: np.ndarray, mean_b: np.ndarray, var_b: np.ndarray):
        """Fold (count, mean, population variance) per feature row into the running statistics."""
        f = n_b.shape[0]
        if f > self.num_samples.shape[0]:
            grow = f - self.num_samples.shape[0]
            self.num_samples = np.append(self.num_samples, np.zeros((grow, 1), dtype=np.int64), axis=0)
            self._mean = np.append(self._mean, np.zeros((grow, 1), dtype=np.float64), axis=0)
            self._variance = np.append(self._variance, np.zeros((grow, 1), dtype=np.float64), axis=0)
        n_a = self.num_samples[:f]
        sum_ns = n_a + n_b
        prod_ns = n_a * n_b
        safe = np.maximum(sum_ns, 1)
        delta2 = (mean_b - self._mean[:f]) ** 2.0
        new_mean = (self._mean[:f] * n_a + mean_b * n_b) / safe
        new_var = (var_b * n_b + n_a * self._variance[:f] + delta2 * prod_ns / safe) / safe
        self._mean[:f] = new_mean
        self._variance[:f] = new_var
        self.num_samples[:f] = sum_ns

    def merge(self, other: "RunningStatsVec"):
        self.merge_triple(other.num_samples, other._mean, other._variance)

    def all_reduce(self, group=None):
        """Combine the statistics of all ranks (every rank ends with the same global triple)."""
        from ..dist import allgather_triples

        triples = allgather_triples(self.num_samples, self._mean, self._variance, group)
        self.clear()
        for n_b, mean_b, var_b in triples:
            self.merge_triple(n_b, mean_b, var_b)
        return self

    @property
    def mean(self):
        return self._mean if self.num_samples.sum() else 0.0

    @property
    def variance(self):
        return self._variance if self.num_samples.sum() else 0.0

    @property
    def std(self):
        return np.sqrt(self._variance) if self.num_samples.sum() else 0.0
