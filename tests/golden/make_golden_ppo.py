"""Generate golden vectors from the reference's own torch/numpy code (build container only).

    python tests/golden/make_golden_ppo.py

The jax-free reference modules are loaded BY FILE PATH (importing the ``src.ppo`` package
would pull in jax through ``ppo_trainer`` -> ``src/runs/batch_runner.py``):
  src/ppo/data_loader.py       PPODataset._compute_gae_returns (:103-130), normalisation (:61-67)
  src/ppo/rollout_buffer.py    RolloutBuffer.store_batch / get_buffer_data (:128-206)
  src/stats/running_stats_vec.py  RunningStatsVec.push (:32-87)
  src/ppo/ppo_agent.py (+ transformer_encoder.py)  masking (:117-121), Categorical log-prob (:182-189)

Output: tests/golden/ppo_reference.npz (inputs and the reference's outputs, all small).
"""
import importlib.util
import sys
import types
from pathlib import Path

import numpy as np
import torch

REF = Path("/root/reference/src")
OUT = Path(__file__).resolve().parent / "ppo_reference.npz"


def load(name: str, path: Path, package: str | None = None):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def main() -> int:
    data_loader = load("ref_data_loader", REF / "ppo/data_loader.py")
    rollout_buffer = load("ref_rollout_buffer", REF / "ppo/rollout_buffer.py")
    stats = load("ref_running_stats_vec", REF / "stats/running_stats_vec.py")
    # ppo_agent uses a relative import of transformer_encoder: give it a stub package
    for name, path in (("refsrc", REF), ("refsrc.ppo", REF / "ppo")):
        pkg = types.ModuleType(name)
        pkg.__path__ = [str(path)]
        sys.modules[name] = pkg
    load("refsrc.env_definitions", REF / "env_definitions.py")
    load("refsrc.ppo.transformer_encoder", REF / "ppo/transformer_encoder.py")
    ppo_agent = load("refsrc.ppo.ppo_agent", REF / "ppo/ppo_agent.py")

    rng = np.random.default_rng(2048)
    out = {}

    # ---- GAE known answers (the reference has none in tests/) -------------
    def gae_case(tag, n, gamma, lam, done_rate, end_done=True):
        rewards = (rng.integers(0, 64, n) * 4 * (rng.random(n) < 0.4)).astype(np.float32)
        values = rng.standard_normal(n).astype(np.float32) * 10
        dones = rng.random(n) < done_rate
        if end_done and n:
            dones[-1] = True
        buf = {
            "observations": np.zeros((n, 16, 31), np.float32),
            "actions": np.zeros((n, 4), np.float32),
            "action_masks": np.ones((n, 4), bool),
            "rewards": rewards,
            "values": values,
            "log_probs": np.zeros(n, np.float32),
            "terminations": dones,
        }
        ds = data_loader.PPODataset(buf, gamma=gamma, lambda_gae=lam)
        adv, ret = ds._compute_gae_returns()
        out[f"gae_{tag}_params"] = np.array([gamma, lam], np.float64)
        out[f"gae_{tag}_rewards"] = rewards
        out[f"gae_{tag}_values"] = values
        out[f"gae_{tag}_dones"] = dones
        out[f"gae_{tag}_adv"] = adv.numpy()
        out[f"gae_{tag}_ret"] = ret.numpy()
        out[f"gae_{tag}_adv_norm"] = ds.advantages.numpy()
        out[f"gae_{tag}_ret_norm"] = ds.returns.numpy()

    gae_case("default", 6000, 0.99, 0.95, 1 / 300)
    gae_case("short_eps", 3000, 0.99, 0.95, 1 / 7)
    gae_case("undiscounted", 2000, 1.0, 1.0, 1 / 100)
    gae_case("lowlam", 2000, 0.9, 0.5, 1 / 50)
    gae_case("open_tail", 1500, 0.99, 0.95, 1 / 200, end_done=False)
    gae_case("two", 2, 0.99, 0.95, 1.0)

    # ---- RolloutBuffer compaction ---------------------------------------------------
    b, t = 7, 11
    boards = rng.integers(0, 12, (b, t, 16))
    obs = np.zeros((b, t, 16, 31), bool)
    np.put_along_axis(obs, boards[..., None], True, axis=-1)
    obs = obs.reshape(b, t, 4, 4, 31)
    acts = rng.integers(0, 4, (b, t))
    acts_onehot = np.eye(4, dtype=np.float32)[acts]
    masks = rng.random((b, t, 4)) < 0.7
    rew = (rng.integers(0, 32, (b, t)) * 4).astype(np.float32)
    val = rng.standard_normal((b, t)).astype(np.float32)
    logp = -rng.random((b, t)).astype(np.float32)
    term = np.zeros((b, t), bool)
    term[0, 4] = True
    term[1, 10] = True
    term[2, 0] = True
    term[3, [3, 7, 9]] = True  # several: first wins
    # env 4: never terminates -> nothing stored
    term[5, 10] = True
    term[6, 6:] = True  # frozen tail as the runner produces it
    rb = rollout_buffer.RolloutBuffer(observation_dim=31, observation_length=16, action_dim=4)
    rb.store_batch(obs, acts_onehot, masks, rew, val, logp, term)
    rb.store_batch(obs[:3], acts_onehot[:3], masks[:3], rew[:3], val[:3], logp[:3], term[:3])
    got = rb.get_buffer_data()
    out.update(
        rb_boards=boards.astype(np.uint8), rb_actions=acts.astype(np.int32), rb_masks=masks,
        rb_rewards=rew, rb_values=val, rb_log_probs=logp, rb_terminations=term,
        rb_size=np.array([rb.buffer_size]),
    )
    for k, v in got.items():
        out[f"rb_out_{k}"] = v

    # ---- RunningStatsVec ------------------------------------------------------------
    rs = stats.RunningStatsVec()
    pushes = [rng.standard_normal((3, n)) * s + m for n, s, m in ((100, 1, 0), (37, 5, 10), (1000, 0.1, -3), (1, 1, 1))]
    for i, p in enumerate(pushes):
        rs.push(p)
        out[f"rs_push{i}"] = p
        out[f"rs_mean{i}"] = rs.mean.copy()
        out[f"rs_var{i}"] = rs.variance.copy()
        out[f"rs_n{i}"] = rs.num_samples.copy()

    # ---- masked logits / Categorical log-prob / entropy -----------------------------
    torch.manual_seed(7)
    agent = ppo_agent.PPOAgent(
        observation_dim=31, action_dim=4, d_model=32, nhead=4, num_layers=1,
        dim_feedforward=64, hidden_dim=32, dropout=0.0,
    ).eval()
    n = 256
    bd = rng.integers(0, 12, (n, 16))
    ob = np.zeros((n, 16, 31), np.float32)
    np.put_along_axis(ob, bd[..., None], 1.0, axis=-1)
    mk = rng.random((n, 4)) < 0.6
    mk[mk.sum(1) == 0, 0] = True
    mk[:4] = np.eye(4, dtype=bool)  # single legal action
    with torch.no_grad():
        raw, value = agent(torch.from_numpy(ob), None)
        masked, _ = agent(torch.from_numpy(ob), torch.from_numpy(mk))
        torch.manual_seed(11)
        actions = torch.distributions.Categorical(logits=masked).sample()
        logp_eval, _, entropy = agent.evaluate_actions(torch.from_numpy(ob), actions, torch.from_numpy(mk))
    # spread the logits so that the fixture also covers large magnitudes
    out.update(
        lp_raw_logits=raw.numpy(), lp_values=value.numpy(), lp_masks=mk, lp_masked_logits=masked.numpy(),
        lp_actions=actions.numpy().astype(np.int32), lp_log_probs=logp_eval.numpy(), lp_entropy=entropy.numpy(),
    )
    big = torch.from_numpy((rng.standard_normal((n, 4)) * 8).astype(np.float32))
    bigm = big - 1e8 * (1 - torch.from_numpy(mk).float())
    torch.manual_seed(13)
    dist = torch.distributions.Categorical(logits=bigm)
    a2 = dist.sample()
    out.update(
        lp2_raw_logits=big.numpy(), lp2_masked_logits=bigm.numpy(), lp2_actions=a2.numpy().astype(np.int32),
        lp2_log_probs=dist.log_prob(a2).numpy(), lp2_entropy=dist.entropy().numpy(),
    )

    np.savez_compressed(OUT, **out)
    print("wrote", OUT, f"{OUT.stat().st_size/1024:.1f} KiB", len(out), "arrays")
    return 0


if __name__ == "__main__":
    sys.exit(main())
