"""Does an L2 flush (a large fill) before the launch change the embedding forward's time?  Prints every rep."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "2048-ppo-agent_b200")]
import torch
from g2048 import engine as E

n = 1 << 18
boards = torch.randint(0, 1 << 62, (n,), dtype=torch.int64, device="cuda")
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
for dt in (torch.float32, torch.bfloat16):
    table = torch.randn(31, 256, device="cuda").to(dt)
    out = torch.empty((n, 16, 256), dtype=dt, device="cuda")
    for entry in ("g2048_embed_boards_plain", "g2048_embed_boards_bulk"):
        for do_flush in (False, True):
            E.embed_boards(boards, table, out=out, entry=entry)
            ts = []
            for _ in range(8):
                if do_flush:
                    flush.fill_(3)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); E.embed_boards(boards, table, out=out, entry=entry); b.record(); torch.cuda.synchronize()
                ts.append(round(a.elapsed_time(b) * 1e3))
            print(str(dt).split(".")[-1], entry[13:], "flush" if do_flush else "no-flush", ts)
