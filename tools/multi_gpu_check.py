"""Functional check of the sharded path on real GPUs over NCCL (torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/multi_gpu_check.py

Every rank runs its shard of the same batches with global env indices; rank 0 also runs the whole batches alone.
Checked: the all-reduced play statistics equal the single-GPU statistics, the gathered per-env shards equal the
single-GPU arrays, recorded rollouts (built-in and network policy, also with live-env compaction) agree shard by
shard, the recording form's flat buffers and sample records concatenate to the single-GPU ones, the key chains stay in
step, and GAE normalisation with all-reduced moments equals the unsharded one.
"""
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "2048-ppo-agent_b200")]
import torch
import torch.distributed as dist

import g2048
from g2048 import dist as gd
from g2048 import engine as E
from g2048.ppo import compute_gae


class RowwiseAgent(torch.nn.Module):
    def __init__(self):
        super().__init__()
        g = torch.Generator().manual_seed(11)
        self.w = torch.nn.Parameter(torch.randn(5, 496, generator=g) * 0.3)

    def forward(self, obs, mask=None):
        out = (obs.reshape(obs.shape[0], 1, 496) * self.w.unsqueeze(0)).sum(dim=2)
        return out[:, :4], out[:, 4:5]


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl")
    dev = torch.device("cuda", torch.cuda.current_device())
    checks = []

    def gather(t):
        parts = [torch.empty_like(t) for _ in range(world)] if t.shape[0] * world else None
        sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([t.shape[0]], dtype=torch.int64, device=dev))
        out = []
        for r in range(world):
            buf = torch.empty((int(sizes[r]),) + tuple(t.shape[1:]), dtype=t.dtype, device=dev)
            if r == rank:
                buf.copy_(t)
            dist.broadcast(buf, src=r)
            out.append(buf)
        return torch.cat(out)

    n = 100_003  # odd on purpose: ragged shards
    for policy, act in (("random", g2048.act_randomly), ("drul", g2048.act_drul)):
        sharded = g2048.BatchRunner(init_seed=17, act_fn=act, shard=(rank, world))
        single = g2048.BatchRunner(init_seed=17, act_fn=act) if rank == 0 else None
        for _ in range(2):
            out = sharded.run_stats_batch(n, per_env=True)
            boards, lengths = gather(out["final_boards"]), gather(out["lengths"])
            if rank == 0:
                ref = single.run_stats_batch(n, per_env=True)
                checks.append((f"{policy}: statistics", ref["summary"] == out["summary"]))
                checks.append((f"{policy}: final boards", torch.equal(boards, ref["final_boards"])))
                checks.append((f"{policy}: lengths", torch.equal(lengths, ref["lengths"])))
                checks.append((f"{policy}: key chain", bool((single.key == sharded.key).all())))

    # the recording form, sharded: every rank records its own envs; shard after shard the flat buffers ARE the
    # single-GPU flat buffer (shards are contiguous env ranges), and so are the 32-byte sample records built from it
    # once the normalisation moments are all-reduced
    for policy, act in (("random", g2048.act_randomly), ("drul", g2048.act_drul)):
        sharded = g2048.BatchRunner(init_seed=23, act_fn=act, shard=(rank, world))
        fr = sharded.run_flat_batch(20_001)
        buf = g2048.RolloutBuffer(31, 16, 4)
        buf.store_flat(fr)
        batches = g2048.DevicePPOBatches(buf.get_packed(), 0.99, 0.95, batch_size=256, group=dist.group.WORLD)
        fields = {k: gather(getattr(fr, k)) for k in ("boards", "meta", "rewards", "log_probs", "lengths", "final_boards")}
        records = gather(batches.records)
        if rank == 0:
            single = g2048.BatchRunner(init_seed=23, act_fn=act)
            ref = single.run_flat_batch(20_001)
            checks.append((f"{policy}: recorded flat buffer", all(torch.equal(fields[k], getattr(ref, k)) for k in fields)))
            checks.append((f"{policy}: recorded summary", ref.summary == fr.summary))
            checks.append((f"{policy}: key chain after recording", bool((single.key == sharded.key).all())))
            ref_buf = g2048.RolloutBuffer(31, 16, 4)
            ref_buf.store_flat(ref)
            ref_rec = g2048.DevicePPOBatches(ref_buf.get_packed(), 0.99, 0.95, batch_size=256).records
            same_ints = torch.equal(records[:, :3], ref_rec[:, :3])  # board, meta + reward, log-prob + value: exact
            floats = records[:, 3:].contiguous().view(torch.float32)  # normalised advantage and return
            ref_floats = ref_rec[:, 3:].contiguous().view(torch.float32)
            checks.append((f"{policy}: sample records of the sharded buffer", same_ints and torch.allclose(floats, ref_floats, rtol=1e-5, atol=1e-6)))

    fn = g2048.TorchActionFunction(RowwiseAgent(), use_mask=True, device=dev)
    for compact in (False, True):
        sharded = g2048.BatchRunner(init_seed=3, act_fn=fn, shard=(rank, world), compact_live=compact)
        ro = sharded.run_packed_batch(501)
        t_all = gd.allreduce_max_int(ro.t_steps, dev)
        meta = gather(ro.meta.t().contiguous())  # (n_shard, T) -> (n, T)
        finals = gather(ro.final_boards)
        if rank == 0:
            single = g2048.BatchRunner(init_seed=3, act_fn=fn)
            ref = single.run_packed_batch(501)
            live = torch.arange(ref.t_steps, device=dev).unsqueeze(0) < ref.lengths().long().unsqueeze(1)
            checks.append((f"network policy (compact={compact}): loop steps", t_all == ref.t_steps == ro.t_steps))
            checks.append((f"network policy (compact={compact}): final boards", torch.equal(finals, ref.final_boards)))
            checks.append((f"network policy (compact={compact}): live records", torch.equal(meta[live], ref.meta.t()[live])))
            checks.append((f"network policy (compact={compact}): key chain", bool((single.key == sharded.key).all())))

    # GAE normalisation on a sharded buffer: the moments are all-reduced, the result equals the unsharded one
    torch.manual_seed(5)
    total = 400_000
    r = (torch.randint(0, 32, (total,), device=dev) * 4).float()
    v = torch.randn(total, device=dev)
    d = (torch.rand(total, device=dev) < 1 / 200).to(torch.uint8)
    d[-1] = 1
    for t in (r, v, d):
        dist.broadcast(t, src=0)
    ends = torch.nonzero(d).flatten()
    cut = int(ends[len(ends) * (rank + 1) // world - 1]) + 1 if rank < world - 1 else total
    lo = 0 if rank == 0 else int(ends[len(ends) * rank // world - 1]) + 1
    adv, ret = compute_gae(r[lo:cut].contiguous(), v[lo:cut].contiguous(), d[lo:cut].contiguous(), 0.99, 0.95, normalize=True,
                           group=dist.group.WORLD)
    adv_all = gather(adv)
    if rank == 0:
        ref_adv, _, mom = E.gae_flat(r, v, d, 0.99, 0.95)  # engine calls: no collective on this solo run
        E.normalize_(ref_adv, mom, 1)
        checks.append(("sharded GAE + all-reduced normalisation", torch.allclose(adv_all, ref_adv, rtol=1e-6, atol=1e-6)))

    if rank == 0:
        for name, ok in checks:
            print(("ok   " if ok else "FAIL ") + name)
        print("multi-GPU check:", "all passed" if all(ok for _, ok in checks) else "FAILED", f"({world} GPUs, NCCL)")
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and not all(ok for _, ok in checks):
        sys.exit(1)


if __name__ == "__main__":
    main()
