"""One launch of the flat GAE kernel at 2^24 steps for ncu source-level capture."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "2048-ppo-agent_b200")]
import torch
from g2048 import engine as E
ng = 1 << 24
r = torch.rand(ng, device="cuda"); v = torch.rand(ng, device="cuda"); d = (torch.rand(ng, device="cuda") < 1 / 300).to(torch.uint8)
for _ in range(3):
    E.gae_flat(r, v, d, 0.99, 0.95)
torch.cuda.synchronize()
print("done")
