"""Ad-hoc timing of the main kernels (development aid; bench.py is the contract)."""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "2048-ppo-agent_b200")]
import numpy as np, torch
from g2048 import engine as E

def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) * 1e-3)
    return best

sms = 148
sink = torch.empty(sms * 8 * 256, dtype=torch.int32, device="cuda")
iters = 20000
t = timed(lambda: E.int_peak_probe(sms * 8, 256, iters, sink))
ops = sms * 8 * 256 * iters * 16 * 3
print(f"int probe: {ops / t / 1e12:.2f} T int-instr/s  ({t*1e3:.2f} ms)")

for mode in (1, 0):
    key = E.words_tensor([0, 7], "cuda")
    subs = E.chain_advance(key, mode, 1 + 2 * 4096)
    t_chain = timed(lambda: E.chain_advance(E.words_tensor([0, 7], "cuda"), mode, 1 + 2 * 4096))
    print(f"mode {mode}: chain of 8193 subs {t_chain*1e3:.2f} ms")
    for policy, name in ((0, "random"), (1, "drul")):
        for n in (1 << 20, 1 << 22):
            stats = None
            def run():
                global stats
                stats = E.play(policy, subs, n, 0, n, mode, per_env=False)["stats"]
            t = timed(run, reps=2)
            st = E.play_stats_dict(stats)
            print(f"mode {mode} {name} n={n}: {st['env_steps'] / t / 1e9:.3f} G env-steps/s  ({t*1e3:.1f} ms, mean len {st['env_steps']/n:.1f}, longest {st['longest']})")

n = 1 << 24
r = torch.rand(n, device="cuda"); v = torch.rand(n, device="cuda"); d = (torch.rand(n, device="cuda") < 1 / 300).to(torch.uint8)
t = timed(lambda: E.gae_flat(r, v, d, 0.99, 0.95))
print(f"gae_flat n={n}: {n * 17 / t / 1e9:.0f} GB/s ({t*1e6:.0f} us) [includes scratch/alloc]")
T, B = 128, 1 << 16
rr = torch.rand(T, B, device="cuda"); vv = torch.rand(T, B, device="cuda"); mm = (torch.rand(T, B, device="cuda") < 1 / 300).to(torch.uint8) << 6
t = timed(lambda: E.gae_time_major(rr, vv, mm, T, B, None, 0.99, 0.95))
print(f"gae_time_major {T}x{B}: {T * B * 17 / t / 1e9:.0f} GB/s ({t*1e6:.0f} us)")
boards = torch.randint(0, 2**62, (1 << 20,), device="cuda")
for dt, sz in ((torch.float32, 4), (torch.bfloat16, 2), (torch.bool, 1)):
    out = torch.empty((1 << 20, 16, 31), dtype=dt, device="cuda")
    t = timed(lambda: E.expand_obs(boards, dt, out=out))
    print(f"expand_obs {dt}: {(1 << 20) * (496 * sz + 8) / t / 1e9:.0f} GB/s ({t*1e6:.0f} us)")
