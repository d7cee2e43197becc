"""How much of a small HBM-bound launch's time is the L2 flush that precedes it?  A 512 MiB fill leaves ~L2-size worth of
DIRTY lines behind; the timed kernel then shares DRAM with their write-back.  Times gae_time_major (C3, 143 MB) and the
65 536-sample gather (134 MB) three ways: (a) right after the write flush, (b) write flush, then a read pass over another
buffer larger than L2 (clean lines only are left), (c) no flush at all, rotating over buffer sets whose combined
footprint is > 3 x L2, back to back (steady state: every launch pays for its OWN write-backs and nobody else's).
    python tools/probes/flush_probe.py"""
import json
import statistics
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "2048-ppo-agent_b200")]

import torch

from g2048 import _native as N
from g2048 import engine as E

dev = torch.device("cuda", 0)
flush_buf = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
drain_buf = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
drain_buf.fill_(1)


def ev():
    return torch.cuda.Event(enable_timing=True)


def timed_single(fn, mode, reps=7):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        if mode in ("write", "write+read"):
            flush_buf.fill_(3)
        if mode == "write+read":
            drain_buf.sum()
        a, b = ev(), ev()
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return statistics.median(ts)


def timed_rotating(fns, passes=7):
    for f in fns:
        f()
    torch.cuda.synchronize()
    ts = []
    for _ in range(passes):
        a, b = ev(), ev()
        a.record()
        for f in fns:
            f()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3 / len(fns))
    return statistics.median(ts)


out = {}
t_steps, b = 128, 1 << 16
mom = torch.zeros(6, dtype=torch.float64, device=dev)


def tm_set():
    rr = torch.rand((t_steps, b), device=dev)
    vv = torch.rand((t_steps, b), device=dev)
    mm = ((torch.rand((t_steps, b), device=dev) < 1 / 300).to(torch.uint8) << 6)
    adv = torch.empty((t_steps, b), dtype=torch.float32, device=dev)
    ret = torch.empty((t_steps, b), dtype=torch.float32, device=dev)
    return lambda: N.call("g2048_gae_time_major", N.ptr(rr), N.ptr(vv), N.ptr(mm), t_steps, b, None, 0.99, 0.95,
                          N.ptr(adv), N.ptr(ret), N.ptr(mom), N.stream_ptr())


sets = [tm_set() for _ in range(4)]
out["gae_time_major C3 (143 MB)"] = {m: timed_single(sets[0], m) for m in ("write", "write+read", "none")}
out["gae_time_major C3 (143 MB)"]["rotating x4"] = timed_rotating(sets)
del sets

n_buf, m = 1 << 22, 1 << 16
records = torch.randint(0, 1 << 62, (n_buf, 4), dtype=torch.int64, device=dev)


def gather_set(count):
    idx = torch.randint(0, n_buf, (count,), device=dev)
    outb = E.minibatch_buffers(count, dev)
    return lambda: E.gather_samples(idx, records, out=outb)


sets = [gather_set(m) for _ in range(4)]
out["gather_samples 65536 (134 MB)"] = {k: timed_single(sets[0], k) for k in ("write", "write+read", "none")}
out["gather_samples 65536 (134 MB)"]["rotating x4"] = timed_rotating(sets)
del sets
big = gather_set(1 << 19)
out["gather_samples 2^19 (1.07 GB)"] = {k: timed_single(big, k) for k in ("write", "write+read", "none")}
del big

n_g = 1 << 26
r, v = torch.rand(n_g, device=dev), torch.rand(n_g, device=dev)
d = (torch.rand(n_g, device=dev) < 1 / 300).to(torch.uint8)
adv, ret = torch.empty(n_g, device=dev), torch.empty(n_g, device=dev)
scr = torch.zeros(int(N.lib.g2048_gae_scan_scratch_bytes(n_g)), dtype=torch.uint8, device=dev)
scan = lambda: N.call("g2048_gae_flat_scan", N.ptr(r), N.ptr(v), N.ptr(d), n_g, 0.99, 0.95, N.ptr(adv), N.ptr(ret), N.ptr(scr),  # noqa: E731
                      N.ptr(mom), N.stream_ptr())
out["gae_scan 2^26 (1.14 GB)"] = {k: timed_single(scan, k) for k in ("write", "write+read", "none")}
print(json.dumps(out, indent=1))
