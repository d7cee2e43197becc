"""Launches of the flat GAE kernels (dispatcher -> pipelined, and the tiled one) at 2^25 steps for ncu source-level capture."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "2048-ppo-agent_b200")]
import torch
from g2048 import engine as E
ng = 1 << 25
r = torch.rand(ng, device="cuda"); v = torch.rand(ng, device="cuda"); d = (torch.rand(ng, device="cuda") < 1 / 300).to(torch.uint8)
for _ in range(2):
    E.gae_flat(r, v, d, 0.99, 0.95)
E.gae_flat(r, v, d, 0.99, 0.95, entry="g2048_gae_flat_tiled")
torch.cuda.synchronize()
print("done")
