"""Throughput of the play kernel against the batch size: how much of the C5 shard's time is the drain at the end
of the launch (queue empty, the last episodes finishing on a few lanes)?"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "2048-ppo-agent_b200")]
import torch
from g2048 import engine as E

key = E.words_tensor([0, 7], "cuda")
subs = E.chain_advance(key, 1, 1 + 2 * 4096)
for policy, name in ((0, "random"), (1, "drul")):
    for lg in (17, 19, 21, 23, 25):
        n = 1 << lg
        out = {}
        def run():
            out.update(E.play(policy, subs, n, 0, n, 1, per_env=False))
        run(); torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); run(); b.record(); torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b) * 1e-3)
        st = E.play_stats_dict(out["stats"])
        print(f"{name:6s} n=2^{lg}: {st['env_steps'] / best / 1e9:7.3f} G env-steps/s  {best * 1e3:8.2f} ms  longest {st['longest']}")
