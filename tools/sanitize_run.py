"""Target of compute-sanitizer (one tool per gpurun call, B200_PROFILING.md): every kernel and C entry once at small,
ragged sizes (tests/kernel_exercise.py), results synchronised.

    compute-sanitizer --tool memcheck  --error-exitcode 1 python tools/sanitize_run.py
    compute-sanitizer --tool racecheck --error-exitcode 1 python tools/sanitize_run.py
"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "2048-ppo-agent_b200"), str(ROOT / "tests")]

import torch

import g2048
import kernel_exercise
from g2048 import _native as N
from g2048 import engine as E

legacy = ROOT / "tests" / "legacy" / "libg2048_legacy.so"
if legacy.exists():
    N.register_entry_points(legacy, N.LEGACY_SIGNATURES)
kernel_exercise.run_all(E, g2048)
# the shared-memory table kernel and its recording form with enough envs for lanes to take several episodes
subs = E.chain_advance(E.words_tensor([0, 5], "cuda"), 1, 1 + 2 * 1024)
for policy in (0, 1):
    E.play(policy, subs, 5000, 0, 5000, 1, entry="g2048_play_tables")
    rec = E.play_record(policy, subs, 5000, 0, 5000, 1)
    offs = E.exclusive_scan(rec["lengths"])
    E.play_record_compact(rec, offs, int(offs[-1]))
# flat GAE kernels above their tile sizes (several tiles per CTA, look-back between tiles)
n = 70_001
r, v = torch.rand(n, device="cuda"), torch.rand(n, device="cuda")
d = (torch.rand(n, device="cuda") < 0.003).to(torch.uint8)
for entry in ("g2048_gae_flat_pipelined", "g2048_gae_flat_tiled", "g2048_gae_flat_scan"):
    E.gae_flat(r, v, d, 0.99, 0.95, entry=entry)
torch.cuda.synchronize()
print("sanitize_run: done")
