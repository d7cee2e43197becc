"""One launch of the recording play kernel's compaction at C4 size (2^18 envs) for an `ncu --set full` capture:
    python tools/profile_compact.py && ncu --set full --clock-control none --import-source on -k regex:play_record_compact \
        -o gpurun_out/r02_compact python tools/profile_compact.py"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "2048-ppo-agent_b200")]

import torch

from g2048 import engine as E

dev = torch.device("cuda:0")
mode = E.RNG_PARTITIONABLE
subs = E.chain_advance(E.words_tensor([0, 4], dev), mode, 1 + 2 * 2048)
n = 1 << 18
rec = E.play_record(E.POLICY_RANDOM, subs, n, 0, n, mode)
offsets = E.exclusive_scan(rec["lengths"])
total = int(offsets[-1])
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
flush.fill_(3)
torch.cuda.synchronize()
flat = E.play_record_compact(rec, offsets, total)
torch.cuda.synchronize()
print("steps", total)
