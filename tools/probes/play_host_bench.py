"""g2048_play_host with pageable (numpy) and pinned result arrays, 2^21 envs, random policy."""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "2048-ppo-agent_b200")]
import numpy as np, torch
from g2048 import _native as N

n = 1 << 21
def bufs(pinned):
    mk = (lambda c, dt: torch.empty(c, dtype=dt, pin_memory=True).numpy()) if pinned else (lambda c, dt: torch.empty(c, dtype=dt).numpy())
    return mk(n, torch.int64).view(np.uint64), mk(n, torch.int32).view(np.uint32), mk(n, torch.int32).view(np.uint32)
stats = np.zeros(N.PLAY_STATS_WORDS, np.uint64)
for pinned in (False, True):
    b, l, s = bufs(pinned)
    ts = []
    for i in range(5):
        t0 = time.perf_counter()
        N.call("g2048_play_host", 0, 2048 + i, None, n, 0, n, 1, b.ctypes.data, l.ctypes.data, s.ctypes.data, stats.ctypes.data)
        ts.append(time.perf_counter() - t0)
    assert int(l.sum()) == int(stats[1])
    print("pinned" if pinned else "pageable", "result arrays:", round(min(ts[1:]) * 1e3, 2), "ms per call,", round(int(stats[1]) / min(ts[1:]) / 1e9, 2), "G env-steps/s")
