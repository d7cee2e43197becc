"""g2048_play at small batch sizes (the SWAR kernel below 32 768 envs) and the single-env replay: shipped library against a
build with -DG2048_PLAY2_NARROW=0 (tools/ab/libg2048_noahead.so: 256-thread CTAs at every size).  python tools/probes/small_batch_probe.py"""
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
dev = torch.device("cuda:0")
subs = E.chain_advance(E.words_tensor([0, 5], dev), 1, 1 + 2 * 2048)
for policy in (0, 1):
    row = []
    for n in (32, 256, 1024, 4096, 16384, 30000):
        E.play(policy, subs, n, 0, n, 1)
        torch.cuda.synchronize()
        ts = []
        for _ in range(7):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); out = E.play(policy, subs, n, 0, n, 1, per_env=False); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        st = E.play_stats_dict(out["stats"])
        row.append("%%d envs %%.3f ms (longest %%d)" %% (n, min(ts), st["longest"]))
    print("policy", policy, "; ".join(row))
""" % (str(ROOT), str(ROOT / "2048-ppo-agent_b200"))
for lib in (None, ROOT / "tools" / "ab" / "libg2048_noahead.so"):
    env = dict(os.environ) if lib is None else dict(os.environ, G2048_LIB=str(lib))
    res = subprocess.run([sys.executable, "-c", CODE], env=env, capture_output=True, text=True)
    print("shipped" if lib is None else lib.name)
    print(res.stdout.strip() or res.stderr[-600:])
