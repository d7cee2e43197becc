"""Minibatch gather timing (the bench.py rows, alone): python tools/probes/gather_bench.py"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path[:0] = [str(ROOT), str(ROOT / "2048-ppo-agent_b200")]

import torch

from g2048 import engine as E


def main():
    dev = torch.device("cuda:0")
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    n_buf = 1 << 22
    packed = dict(boards=torch.randint(0, 1 << 62, (n_buf,), dtype=torch.int64, device=dev),
                  meta=torch.randint(0, 127, (n_buf,), dtype=torch.uint8, device=dev),
                  log_probs=torch.rand(n_buf, device=dev), values=torch.rand(n_buf, device=dev))
    adv, ret = torch.rand(n_buf, device=dev), torch.rand(n_buf, device=dev)
    for m in (1 << 11, 1 << 16, 1 << 19):
        for dt in (torch.float32, torch.bfloat16, None):
            idx = torch.randint(0, n_buf, (m,), device=dev)
            out = E.minibatch_buffers(m, dev, dt)
            fn = lambda: E.gather_minibatch(idx, packed, adv, ret, dt, out=out)
            fn()
            torch.cuda.synchronize()
            ts = []
            for _ in range(7):
                flush.fill_(1)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); fn(); b.record(); torch.cuda.synchronize()
                ts.append(a.elapsed_time(b) * 1e3)
            ts.sort()
            item = 0 if dt is None else torch.empty(0, dtype=dt).element_size()
            nbytes = m * (8 + 8 + 1 + 16 + 496 * item + 8 + 4 + 16)
            print(f"m={m:7d} obs={str(dt):15s} {ts[3]:8.1f} us  {nbytes / ts[3] / 1e6:6.2f} TB/s")


if __name__ == "__main__":
    main()
