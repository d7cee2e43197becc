"""gae_time_major at C3 (128 x 65 536) and 4 x C3: bulk-copy ring kernel vs one lane per env (G2048_GAE_TM_RING=0), timed
after the bench's write flush and rotating over four buffer sets (steady state).  python tools/probes/tm_ring_probe.py"""
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent.parent
CODE = r"""
import sys, statistics
sys.path[:0] = [%r, %r]
import torch
from g2048 import _native as N
dev = torch.device("cuda:0")
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
mom = torch.zeros(6, dtype=torch.float64, device=dev)
def ev(): return torch.cuda.Event(enable_timing=True)
for b in (1 << 16, 1 << 18):
    T = 128
    sets = []
    for _ in range(4):
        rr, vv = torch.rand((T, b), device=dev), torch.rand((T, b), device=dev)
        mm = ((torch.rand((T, b), device=dev) < 1 / 300).to(torch.uint8) << 6)
        a, r = torch.empty((T, b), device=dev), torch.empty((T, b), device=dev)
        sets.append((rr, vv, mm, a, r))
    def run(s):
        N.call("g2048_gae_time_major", N.ptr(s[0]), N.ptr(s[1]), N.ptr(s[2]), T, b, None, 0.99, 0.95, N.ptr(s[3]), N.ptr(s[4]), N.ptr(mom), N.stream_ptr())
    for s in sets: run(s)
    torch.cuda.synchronize()
    ts = []
    for _ in range(7):
        flush.fill_(3); x, y = ev(), ev(); x.record(); run(sets[0]); y.record(); torch.cuda.synchronize(); ts.append(x.elapsed_time(y) * 1e3)
    rot = []
    for _ in range(7):
        x, y = ev(), ev(); x.record()
        for s in sets: run(s)
        y.record(); torch.cuda.synchronize(); rot.append(x.elapsed_time(y) * 1e3 / 4)
    print(b, "after write flush %%.1f us, rotating %%.1f us" %% (statistics.median(ts), statistics.median(rot)))
""" % (str(ROOT), str(ROOT / "2048-ppo-agent_b200"))
for ring in ("1", "0"):
    res = subprocess.run([sys.executable, "-c", CODE], env=dict(os.environ, G2048_GAE_TM_RING=ring), capture_output=True, text=True)
    print("G2048_GAE_TM_RING=" + ring)
    print(res.stdout.strip() or res.stderr[-600:])
# other ring geometries: builds made with `make OUT=../../tools/ab/libg2048_tm_<envs>_<rows>.so BUILD=build_tm_<envs>_<rows>
# EXTRA="-DG2048_TM_RING_ENVS=<envs> -DG2048_TM_RING_ROWS=<rows>"`
for lib in sorted((ROOT / "tools" / "ab").glob("libg2048_tm_*.so")):
    res = subprocess.run([sys.executable, "-c", CODE], env=dict(os.environ, G2048_LIB=str(lib)), capture_output=True, text=True)
    print(lib.name)
    print(res.stdout.strip() or res.stderr[-600:])
