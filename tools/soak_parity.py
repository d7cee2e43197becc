"""Soak parity runs against the oracle's C port, many seeds, both policies, both Threefry layouts:

  play      whole batches of the persistent play kernel: every final board, length and score;
  recorded  the recording form (run_flat_batch): every recorded step of a window of envs -- the action the policy would
            take (oracle's act on the oracle's own state and keys), the pre-step board, the reward and the done flag the
            oracle's env.step produces for it, and the legal mask;
  policy    g2048_policy_step_obs with random logits and auto-reset: the env transition for the kernel's own sampled
            action (board, mask, done, reward, auto-reset board) and the observation tensor, every step.

python tools/soak_parity.py [--envs 262144] [--seeds 6] [--legs play,recorded,policy]   (exit code 1 on any mismatch)
"""
import argparse
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "2048-ppo-agent_b200")]
import numpy as np
import torch

import g2048
from g2048 import engine as E
from oracle import c_oracle as CO
from oracle import pgx2048_oracle as O


def mask_bits(x):
    return ((x[:, None] >> np.arange(4)) & 1).astype(np.uint8)


def soak_play(envs, seeds, log):
    steps = bad = 0
    for mode in (1, 0):
        for policy in (0, 1):
            for seed in range(seeds):
                s = 1000 * seed + 17 * policy + mode
                subs = E.chain_advance(E.words_tensor(list(E.key_words(s)), "cuda"), mode, 1 + 2 * 2048)
                out = E.play(policy, subs, envs, 0, envs, mode, per_env=True)
                want = CO.play(s, envs, policy, mode, max_steps=2048)
                same = (np.array_equal(E.boards_numpy(out["final_boards"]), want["final_boards"])
                        and np.array_equal(out["lengths"].cpu().numpy(), want["lengths"])
                        and np.array_equal(out["scores"].cpu().numpy(), want["scores"]))
                n = int(want["lengths"].sum())
                steps += n
                bad += 0 if same else 1
                log(f"play      mode {mode} policy {policy} seed {s}: {n} env-steps, {'identical' if same else 'MISMATCH'}")
    return steps, bad


def check_flat_window(flat, chain_key, batch, policy, mode, lo, hi):
    """Every recorded step of envs [lo, hi) of a FlatRollout against the oracle (state, keys and policy of its own)."""
    t_max = int(flat.lengths[lo:hi].max())
    _, subs = CO.chain(np.asarray(chain_key, np.uint32), mode, 1 + 2 * t_max)
    boards, masks = CO.env_init(CO.split(subs[0], batch, mode)[lo:hi], mode)
    n = hi - lo
    done = np.zeros(n, np.uint8)
    offs = flat.offsets[lo:hi + 1].cpu().numpy()
    lens = flat.lengths[lo:hi].cpu().numpy().astype(np.int64)
    a, b = int(offs[0]), int(offs[-1])
    f_boards = E.boards_numpy(flat.boards[a:b])
    f_meta = flat.meta[a:b].cpu().numpy()
    f_rew = flat.rewards[a:b].cpu().numpy()
    ok = True
    for t in range(t_max):
        live = lens > t
        pos = offs[:-1][live] - a + t
        want_a, _ = CO.act(CO.split(subs[1 + 2 * t], batch, mode)[lo:hi], masks, policy, mode)
        ok &= bool(np.array_equal(f_boards[pos], boards[live]))
        ok &= bool(np.array_equal(f_meta[pos] & 3, want_a[live]))
        ok &= bool(np.array_equal(mask_bits((f_meta[pos] >> 2) & 15), masks[live]))
        boards, masks, done, rew = CO.env_step(boards, masks, done, want_a, CO.split(subs[2 + 2 * t], batch, mode)[lo:hi], mode)
        ok &= bool(np.array_equal(rew[live], f_rew[pos]) and np.array_equal(done[live], (f_meta[pos] >> 6) & 1))
        ok &= bool(done[~live].all())
    ok &= bool(done.all()) and bool(np.array_equal(E.boards_numpy(flat.final_boards[lo:hi]), boards))
    return ok, int(lens.sum())


def soak_recorded(envs, seeds, log, window=2048):
    steps = bad = 0
    for mode in (1, 0):
        for policy, fn in ((0, g2048.act_randomly), (1, g2048.act_drul)):
            for seed in range(seeds):
                s = 1000 * seed + 17 * policy + mode + 5
                runner = g2048.BatchRunner(init_seed=s, act_fn=fn, rng_mode=mode)
                key = runner.key
                flat = runner.run_flat_batch(envs)
                lo = (seed * 7919) % max(1, envs - window)
                same, n = check_flat_window(flat, key, envs, policy, mode, lo, min(envs, lo + window))
                steps += n
                bad += 0 if same else 1
                log(f"recorded  mode {mode} policy {policy} seed {s}: {n} recorded steps of envs [{lo},{lo + window}) checked, "
                    f"{'identical' if same else 'MISMATCH'}")
    return steps, bad


def soak_policy(envs, seeds, log, t_steps=48):
    steps = bad = 0
    for mode in (1, 0):
        for seed in range(seeds):
            s = 1000 * seed + mode + 11
            gen = torch.Generator(device="cuda").manual_seed(s)
            subs = E.chain_advance(E.words_tensor(list(E.key_words(s)), "cuda"), mode, 1 + 2 * t_steps)
            boards, status = E.env_init(subs[0], envs, 0, envs, mode)
            obs = E.expand_obs(boards, torch.float32)
            acts = torch.empty(envs, dtype=torch.int32, device="cuda")
            rr = torch.empty(envs, dtype=torch.float32, device="cuda")
            rm = torch.empty(envs, dtype=torch.uint8, device="cuda")
            rb = torch.empty(envs, dtype=torch.int64, device="cuda")
            ob, om = E.boards_numpy(boards), mask_bits(status.cpu().numpy())
            same = True
            for t in range(t_steps):
                logits = torch.randn((envs, 4), device="cuda", generator=gen) * 3
                E.policy_step_obs(boards, status, logits, None, True, True, True, subs[1 + 2 * t:], None, envs, 0, mode, obs,
                                  rb, rm, rr, None, None, acts)
                a = acts.cpu().numpy()
                same &= bool(np.array_equal(E.boards_numpy(rb), ob)) and bool(om[np.arange(envs), a].all())
                keys = CO.split(subs[2 + 2 * t].cpu().numpy().view(np.uint32), envs, mode)
                k0, k1 = O.split((keys[:, 0], keys[:, 1]), 2, mode)
                wb, wm, wd, wr = CO.env_step(ob, om, np.zeros(envs, np.uint8), a, np.stack([k0[:, 0], k1[:, 0]], 1), mode)
                ib, im = CO.env_init(np.stack([k0[:, 1], k1[:, 1]], 1), mode)
                ob = np.where(wd[:, None].astype(bool), ib, wb)
                om = np.where(wd[:, None].astype(bool), im, wm)
                st = status.cpu().numpy()
                same &= bool(np.array_equal(E.boards_numpy(boards), ob) and np.array_equal(mask_bits(st), om)
                             and np.array_equal((st >> 4) & 1, wd) and np.array_equal(rr.cpu().numpy(), wr))
                same &= bool(torch.equal(obs, E.expand_obs(boards, torch.float32)))
            steps += envs * t_steps
            bad += 0 if same else 1
            log(f"policy    mode {mode} seed {s}: {envs * t_steps} env-steps with auto-reset, {'identical' if same else 'MISMATCH'}")
    return steps, bad


LEGS = {"play": soak_play, "recorded": soak_recorded, "policy": soak_policy}


def run(envs, seeds, legs, log=lambda s: print(s, flush=True)):
    CO.set_num_threads(__import__("os").cpu_count() or 1)
    total = bad = 0
    t0 = time.perf_counter()
    for leg in legs:
        n, b = LEGS[leg](envs, seeds, log)
        total += n
        bad += b
    log(f"soak: {total} env-steps over legs {','.join(legs)} ({seeds} seeds, {envs} envs), {bad} mismatching batches, "
        f"{time.perf_counter() - t0:.0f} s")
    return total, bad


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=1 << 18)
    ap.add_argument("--seeds", type=int, default=6)
    ap.add_argument("--legs", default="play,recorded,policy")
    args = ap.parse_args()
    _, bad = run(args.envs, args.seeds, args.legs.split(","))
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
