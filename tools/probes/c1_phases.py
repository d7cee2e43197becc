"""Where the time of the reference-format call run_actions_batch(1024) goes (BASELINE config C1)."""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "2048-ppo-agent_b200")]
import torch
import g2048
from g2048 import engine as E

def t(fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); out = fn(); torch.cuda.synchronize(); return out, (time.perf_counter() - t0) * 1e3

runner = g2048.BatchRunner(init_seed=0, act_fn=g2048.act_randomly)
runner.run_actions_batch(1024)
for rep in range(3):
    ro, t_roll = t(lambda: runner._run(1024, keep_states=False, full_records=True))
    T, b = ro.t_steps, ro.batch_size
    obs, t_obs = t(lambda: E.expand_obs(ro.boards, torch.bool, rows=T, n_cols=b))
    un, t_un = t(lambda: E.unpack_records(ro.meta, ro.rewards, ro.log_probs, ro.values, T, b))
    host, t_d2h = t(lambda: obs.cpu())
    arr, t_np = t(lambda: host.numpy().reshape(b, T, 4, 4, 31))
    rest, t_rest = t(lambda: [None if x is None else x.cpu().numpy() for x in un.values()])
    pin = torch.empty(obs.shape, dtype=obs.dtype, pin_memory=True)
    _, t_pin = t(lambda: pin.copy_(obs, non_blocking=False))
    print(f"T={T}: rollout {t_roll:.2f} ms, expand_obs {t_obs:.2f}, unpack {t_un:.2f}, obs D2H pageable {t_d2h:.2f} ({obs.numel()/1e6:.0f} MB), "
          f"numpy view {t_np:.2f}, other arrays D2H {t_rest:.2f}, obs D2H into existing pinned {t_pin:.2f}")
_, total = t(lambda: runner.run_actions_batch(1024))
print("run_actions_batch total", round(total, 2), "ms")
