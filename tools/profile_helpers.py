"""One launch of each helper kernel that was rewritten at the end of round 2, at C4's size, for an `ncu --set full` capture:
    ncu --set full --clock-control none --import-source on -k regex:'compact_records_kernel|episode_lengths_kernel|scan_chunk' \
        -o gpurun_out/r02_helpers python tools/profile_helpers.py
    python tools/summarize_ncu.py gpurun_out/r02_helpers.ncu-rep profiles/r02_helper_kernels"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "2048-ppo-agent_b200")]

import torch

import g2048
from g2048 import engine as E

dev = torch.device("cuda:0")
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
runner = g2048.BatchRunner(init_seed=4, act_fn=g2048.act_randomly)
ro = runner.run_packed_batch(1 << 18)  # lock-step recorder: (T, B) time-major records
t, b = ro.t_steps, ro.batch_size
flush.fill_(1)
torch.cuda.synchronize()
lengths = E.episode_lengths(ro.meta, t, b)              # episode_lengths_kernel
flush.fill_(1)
torch.cuda.synchronize()
offsets = E.exclusive_scan(lengths)                     # scan_chunk_sums / _bases / scan_chunks
total = int(offsets[-1])
out = dict(boards=torch.empty(total, dtype=torch.int64, device=dev), meta=torch.empty(total, dtype=torch.uint8, device=dev),
           rewards=torch.empty(total, dtype=torch.float32, device=dev), log_probs=torch.empty(total, dtype=torch.float32, device=dev),
           values=torch.zeros(total, dtype=torch.float32, device=dev))
flush.fill_(1)
torch.cuda.synchronize()
E.compact_records(ro.boards, ro.meta, ro.rewards, ro.log_probs, None, t, b, lengths, offsets, 0, out["boards"], out["meta"],
                  out["rewards"], out["log_probs"], None)  # compact_records_kernel
torch.cuda.synchronize()
print("loop steps", t, "kept steps", total)
