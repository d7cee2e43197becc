"""Multi-GPU plumbing: env sharding is free (envs are independent, keys use global env indices);
collectives are needed only for the small reductions around the path:

  * episode statistics (play stats block, RunningStatsVec triples),
  * the advantage / return normalisation moments (src/ppo/data_loader.py:61-67 on a sharded buffer),
  * "all envs done" / longest-episode agreement so that every rank advances the key chain alike,
  * PPO gradients (DDP-style bucketed all-reduce).

All of them go through torch.distributed (NCCL on GPUs; gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def is_dist() -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous slice [lo, hi) of n envs owned by `rank`."""
    return n * rank // world, n * (rank + 1) // world


def _comm_device(t: torch.Tensor) -> torch.device:
    return t.device if dist.get_backend() == "nccl" else torch.device("cpu")


def allreduce_sum_(t: torch.Tensor, group=None) -> torch.Tensor:
    if not is_dist():
        return t
    dev = _comm_device(t)
    buf = t if t.device == dev else t.to(dev)
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    if buf is not t:
        t.copy_(buf)
    return t


def allreduce_play_stats(stats: torch.Tensor, group=None) -> torch.Tensor:
    """Sum the play statistics block over ranks; slot 5 (longest episode) is a max.  ONE collective: the 256-byte blocks
    are all-gathered and reduced locally (a SUM and a MAX all-reduce cost two NCCL launches, ~0.05 ms each, per batch)."""
    if not is_dist():
        return stats
    dev = _comm_device(stats)
    mine = stats.to(dev).contiguous()
    world = dist.get_world_size(group)
    flat = torch.empty(world * mine.shape[0], dtype=mine.dtype, device=dev)  # 1-D: gloo wants the concatenated form
    dist.all_gather_into_tensor(flat, mine, group=group)
    gathered = flat.view(world, mine.shape[0])
    out = gathered.sum(dim=0)
    out[5] = gathered[:, 5].max()
    return out.to(stats.device)


def allreduce_max_int(value: int, device, group=None) -> int:
    if not is_dist():
        return value
    dev = device if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor([value], dtype=torch.int64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return int(t.item())


def all_ranks_true(flag: bool, device, group=None) -> bool:
    if not is_dist():
        return flag
    return allreduce_max_int(0 if flag else 1, device, group) == 0


def allgather_triples(n: np.ndarray, mean: np.ndarray, var: np.ndarray, group=None):
    """Gather (count, mean, variance) per feature row from every rank, in rank order."""
    if not is_dist():
        return [(n, mean, var)]
    world = dist.get_world_size(group)
    gathered = [None] * world
    dist.all_gather_object(gathered, (np.asarray(n), np.asarray(mean), np.asarray(var)), group=group)
    return gathered


def allreduce_gradients(params, group=None, bucket_bytes: int = 32 << 20) -> None:
    """Average .grad over ranks in flat buckets (the PPO learner's data-parallel step)."""
    if not is_dist():
        return
    world = dist.get_world_size(group)
    grads = [p.grad for p in params if p.grad is not None]
    bucket, size = [], 0

    def flush():
        nonlocal bucket, size
        if not bucket:
            return
        flat = torch.cat([g.reshape(-1) for g in bucket])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.div_(world)
        off = 0
        for g in bucket:
            g.copy_(flat[off: off + g.numel()].view_as(g))
            off += g.numel()
        bucket, size = [], 0

    for g in grads:
        bucket.append(g)
        size += g.numel() * g.element_size()
        if size >= bucket_bytes:
            flush()
    flush()
