"""Board-embedding kernels: bulk-copy forward vs the plain-store forward, and the table gradient.
Usage: python tools/bench_embed.py   (prints one JSON object; times are CUDA-event medians)"""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "2048-ppo-agent_b200")]
import torch

from g2048 import engine as E


def timed(fn, reps=20, inner=1):
    """Median seconds per call; inner > 1 queues several calls per event pair so that small launches are
    timed back to back on the device instead of at the host's launch rate."""
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(inner):
            fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e-3 / inner)
    return sorted(ts)[len(ts) // 2]


def main():
    out = {}
    for n in (2048, 1 << 18):
        boards = torch.randint(0, 1 << 62, (n,), dtype=torch.int64, device="cuda")
        inner = 20 if n <= 4096 else 1
        for dt in (torch.float32, torch.bfloat16):
            item = 4 if dt == torch.float32 else 2
            table = torch.randn(31, 256, device="cuda").to(dt)
            res = torch.empty((n, 16, 256), dtype=dt, device="cuda")
            nbytes = n * 16 * 256 * item
            row = {}
            for entry in ("g2048_embed_boards", "g2048_embed_boards_bulk", "g2048_embed_boards_plain"):
                t = timed(lambda: E.embed_boards(boards, table, out=res, entry=entry), inner=inner)
                row[entry] = {"us": t * 1e6, "GB/s": nbytes / t / 1e9}
            t = timed(lambda: E.embed_boards_grad(boards, res), inner=inner)
            row["g2048_embed_boards_grad"] = {"us": t * 1e6, "GB/s": nbytes / t / 1e9}
            if dt == torch.float32 and n <= (1 << 18):
                obs = E.expand_obs(boards, torch.float32)
                w = table.t().contiguous()
                t = timed(lambda: torch.nn.functional.linear(obs, w))
                row["torch_linear_on_one_hot_obs"] = {"us": t * 1e6}
                t = timed(lambda: (E.expand_obs(boards, torch.float32, out=obs), torch.nn.functional.linear(obs, w)))
                row["expand_obs_plus_torch_linear"] = {"us": t * 1e6}
            out[f"n={n},{str(dt).split('.')[-1]}"] = row
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
