"""A/B of the play kernels (v1 vs current) on the same box, plus correctness cross-check."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "2048-ppo-agent_b200")]
import numpy as np, torch
from g2048 import _native as N
from g2048 import engine as E

# the first-generation kernels live in the tests' build of the library (make -C 2048-ppo-agent_b200/csrc legacy)
N.register_entry_points(ROOT / "tests" / "legacy" / "libg2048_legacy.so", N.LEGACY_SIGNATURES)

def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) * 1e-3)
    return best

for mode in (1, 0):
    key = E.words_tensor([0, 7], "cuda")
    subs = E.chain_advance(key, mode, 1 + 2 * 4096)
    for policy, name in ((0, "random"), (1, "drul")):
        n = 1 << 21
        res = {}
        for entry in ("g2048_play_v1", "g2048_play_swar", "g2048_play_tables"):
            out = {}
            def run():
                out.update(E.play(policy, subs, n, 0, n, mode, per_env=True, entry=entry))
            t = timed(run)
            st = E.play_stats_dict(out["stats"])
            res[entry] = out
            print(f"mode {mode} {name:6s} {entry:14s} n={n}: {st['env_steps'] / t / 1e9:7.3f} G env-steps/s ({t*1e3:.2f} ms) longest {st['longest']} ovf {st['overflowed']}")
        a, b = res["g2048_play_swar"], res["g2048_play_tables"]
        same = all(torch.equal(a[k], b[k]) for k in ("final_boards", "lengths", "scores"))
        sa, sb = a["stats"].clone(), b["stats"].clone()
        print("   identical per-env results:", same, " identical stats:", torch.equal(sa, sb))

n = 1 << 26
r = torch.rand(n, device="cuda"); v = torch.rand(n, device="cuda"); d = (torch.rand(n, device="cuda") < 1 / 300).to(torch.uint8)
for rate in (1 / 300, 1 / 30, 1 / 3000):
    d = (torch.rand(n, device="cuda") < rate).to(torch.uint8)
    outs = {}
    for entry in ("g2048_gae_flat_v1", "g2048_gae_flat"):
        t = timed(lambda: outs.__setitem__(entry, E.gae_flat(r, v, d, 0.99, 0.95, entry=entry)))
        print(f"{entry:18s} n=2^26 done rate {rate:.5f}: {n * 17 / t / 1e9:6.0f} GB/s ({t*1e6:.0f} us incl. scratch alloc)")
    print("   identical:", torch.equal(outs["g2048_gae_flat_v1"][0], outs["g2048_gae_flat"][0]), torch.equal(outs["g2048_gae_flat_v1"][1], outs["g2048_gae_flat"][1]))
boards = torch.randint(0, 2**62, (1 << 20,), device="cuda")
for dt, sz in ((torch.float32, 4), (torch.bfloat16, 2), (torch.bool, 1)):
    out = torch.empty((1 << 20, 16, 31), dtype=dt, device="cuda")
    ref = None
    for entry in ("g2048_expand_obs_v1", "g2048_expand_obs"):
        t = timed(lambda: E.expand_obs(boards, dt, out=out, entry=entry))
        same = "" if ref is None else f" identical: {torch.equal(ref.view(torch.uint8), out.view(torch.uint8))}"
        ref = out.clone()
        print(f"{entry:20s} {str(dt):15s}: {(1 << 20) * (496 * sz + 8) / t / 1e9:6.0f} GB/s ({t*1e6:.0f} us){same}")
