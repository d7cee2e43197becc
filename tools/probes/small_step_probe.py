"""The fused policy step (g2048_policy_step_obs) and the record gather at small sizes: shipped library (one warp per CTA below
~19 000 envs / samples) against a build with 256-thread CTAs everywhere (tools/ab/libg2048_wide.so: -DG2048_OBS_NARROW=0
-DG2048_PLAY2_NARROW=0).  python tools/probes/small_step_probe.py"""
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent.parent
CODE = r"""
import sys
sys.path[:0] = [%r, %r]
import torch
from g2048 import engine as E
dev = torch.device("cuda:0"); mode = 1; T = 64
def ev(): return torch.cuda.Event(enable_timing=True)
row = []
for b in (256, 512, 2048, 4096, 16384):
    subs = E.chain_advance(E.words_tensor([0, 3], dev), mode, 1 + 2 * T)
    pb, ps = E.env_init(subs[0], b, 0, b, mode)
    logits, values = torch.randn((b, 4), device=dev), torch.randn(b, device=dev)
    obs = torch.empty((b, 16, 31), dtype=torch.float32, device=dev)
    rec = [torch.empty((T, b), dtype=dt, device=dev) for dt in (torch.int64, torch.uint8, torch.float32, torch.float32, torch.float32)]
    def rollout():
        for k in range(T):
            E.policy_step_obs(pb, ps, logits, values, True, True, True, subs[1 + 2 * k:], None, b, 0, mode, obs, *[r[k] for r in rec])
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side): rollout()
    torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g): rollout()
    ts = []
    for _ in range(5):
        a, e = ev(), ev(); a.record(); g.replay(); e.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(e))
    row.append("%%d envs %%.2f us" %% (b, min(ts) * 1e3 / T))
print("fused step, graph replay:", "; ".join(row))
n_buf = 1 << 22
records = torch.randint(0, 1 << 62, (n_buf, 4), dtype=torch.int64, device=dev)
row = []
for m in (512, 2048, 8192):
    idx = torch.randint(0, n_buf, (m,), device=dev)
    out = E.minibatch_buffers(m, dev)
    E.gather_samples(idx, records, out=out); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(32): E.gather_samples(idx, records, out=out)
    ts = []
    for _ in range(5):
        a, e = ev(), ev(); a.record(); g.replay(); e.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(e))
    row.append("%%d samples %%.2f us" %% (m, min(ts) * 1e3 / 32))
print("record gather, graph replay:", "; ".join(row))
""" % (str(ROOT), str(ROOT / "2048-ppo-agent_b200"))
for lib in (None, ROOT / "tools" / "ab" / "libg2048_wide.so"):
    env = dict(os.environ) if lib is None else dict(os.environ, G2048_LIB=str(lib))
    res = subprocess.run([sys.executable, "-c", CODE], env=env, capture_output=True, text=True)
    print("shipped" if lib is None else lib.name)
    print(res.stdout.strip() or res.stderr[-800:])
