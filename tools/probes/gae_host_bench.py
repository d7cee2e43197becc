"""g2048_gae_host (host numpy buffers in, host numpy buffers out) at C4's buffer size."""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "2048-ppo-agent_b200")]
import numpy as np
from g2048 import engine as E
from oracle import c_oracle as CO

rng = np.random.default_rng(0)
for n in (1_000_000, 31_000_000):
    r = (rng.integers(0, 64, n) * 4).astype(np.float32); v = rng.standard_normal(n).astype(np.float32)
    d = (rng.random(n) < 1 / 300).astype(np.uint8)
    E.gae_host(r, v, d, 0.99, 0.95, True)
    ts = []
    for _ in range(3):
        t0 = time.perf_counter(); adv, ret = E.gae_host(r, v, d, 0.99, 0.95, False); ts.append(time.perf_counter() - t0)
    wa, wr = CO.gae(r[:200_000], v[:200_000], np.concatenate([d[:199_999], [1]]).astype(np.uint8), 0.99, 0.95)
    print(f"n={n}: {min(ts) * 1e3:.1f} ms ({n * 17 / min(ts) / 1e9:.2f} GB/s of host traffic), first call incl. allocations excluded")
