"""One launch of each hot kernel at benchmark size, for `ncu --set full` captures (see profiles/)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "2048-ppo-agent_b200")]
import torch
from g2048 import engine as E

dev = "cuda"
mode = 1
n = 1 << 21
key = E.words_tensor([0, 2048], dev)
subs = E.chain_advance(key, mode, 1 + 2 * 2048)
for policy in (0, 1):
    for _ in range(2):
        out = E.play(policy, subs, n, 0, n, mode, per_env=False)
    print("play", policy, E.play_stats_dict(out["stats"])["env_steps"])
ng = 1 << 24
r = torch.rand(ng, device=dev); v = torch.rand(ng, device=dev); d = (torch.rand(ng, device=dev) < 1 / 300).to(torch.uint8)
for _ in range(2):
    E.gae_flat(r, v, d, 0.99, 0.95)
T, B = 128, 1 << 16
rr = torch.rand(T, B, device=dev); vv = torch.rand(T, B, device=dev); mm = (torch.rand(T, B, device=dev) < 1 / 300).to(torch.uint8) << 6
for _ in range(2):
    E.gae_time_major(rr, vv, mm, T, B, None, 0.99, 0.95)
boards = torch.randint(0, 2**62, (1 << 20,), device=dev)
for dt in (torch.float32, torch.bool):
    for _ in range(2):
        E.expand_obs(boards, dt)
pb, ps = E.env_init(subs[0], B, 0, B, mode)
logits = torch.randn(B, 4, device=dev); values = torch.randn(B, device=dev)
rb = torch.empty(B, dtype=torch.int64, device=dev); rm = torch.empty(B, dtype=torch.uint8, device=dev)
r1, r2, r3 = (torch.empty(B, device=dev) for _ in range(3))
for k in range(4):
    E.policy_step(pb, ps, logits, values, True, True, True, subs[1 + 2 * k], subs[2 + 2 * k], B, 0, mode, rb, rm, r1, r2, r3)
torch.cuda.synchronize()
print("done")
