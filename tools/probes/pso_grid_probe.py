"""C3 fused step (g2048_policy_step_obs, 65 536 envs, float32 observations) against the number of resident CTAs per SM
the launch asks for (G2048_PSO_CTAS_PER_SM, read once per process): python tools/probes/pso_grid_probe.py"""
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
dev = torch.device("cuda:0"); mode = 1; b = 1 << 16; T = 128
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
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); g.replay(); e.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(e))
print("us per step (graph replay):", round(min(ts) * 1e3 / T, 2))
""" % (str(ROOT), str(ROOT / "2048-ppo-agent_b200"))
for ctas in (1, 2, 3):
    res = subprocess.run([sys.executable, "-c", CODE], env=dict(os.environ, G2048_PSO_CTAS_PER_SM=str(ctas)), capture_output=True, text=True)
    print("CTAs per SM", ctas, res.stdout.strip() or res.stderr[-400:])
# an explicit grid (G2048_PSO_GRID): 256 = one CTA per 8 tiles (108 SMs get two CTAs, 40 get one); multiples of 148 spread
# the 2 048 tiles evenly over the SMs
for grid in (148, 222, 256, 296, 370, 444):
    res = subprocess.run([sys.executable, "-c", CODE], env=dict(os.environ, G2048_PSO_GRID=str(grid)), capture_output=True, text=True)
    print("grid", grid, res.stdout.strip() or res.stderr[-400:])
