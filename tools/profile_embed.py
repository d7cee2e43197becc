"""One launch of each board-embedding kernel at 2^18 boards, d_model 256 (for ncu)."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "2048-ppo-agent_b200")]
import torch

from g2048 import engine as E

n = 1 << 18
boards = torch.randint(0, 1 << 62, (n,), dtype=torch.int64, device="cuda")
for dt in (torch.float32, torch.bfloat16):
    table = torch.randn(31, 256, device="cuda").to(dt)
    for entry in ("g2048_embed_boards_bulk", "g2048_embed_boards_plain"):
        out = E.embed_boards(boards, table, entry=entry)
    E.embed_boards_grad(boards, out)
torch.cuda.synchronize()
print("done")
