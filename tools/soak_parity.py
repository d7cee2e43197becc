"""Soak: whole batches of the persistent play kernel against the oracle's C port, many seeds, both policies, both
Threefry layouts -- every final board, length and score.  python tools/soak_parity.py [--envs 262144] [--seeds 6]"""
import argparse
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "2048-ppo-agent_b200")]
import numpy as np
import torch

from g2048 import engine as E
from oracle import c_oracle as CO


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=1 << 18)
    ap.add_argument("--seeds", type=int, default=6)
    args = ap.parse_args()
    CO.set_num_threads(__import__("os").cpu_count() or 1)
    total_steps, bad, t0 = 0, 0, time.perf_counter()
    for mode in (1, 0):
        for policy in (0, 1):
            for seed in range(args.seeds):
                s = 1000 * seed + 17 * policy + mode
                subs = E.chain_advance(E.words_tensor(list(E.key_words(s)), "cuda"), mode, 1 + 2 * 2048)
                out = E.play(policy, subs, args.envs, 0, args.envs, mode, per_env=True)
                want = CO.play(s, args.envs, policy, mode, max_steps=2048)
                same = (np.array_equal(E.boards_numpy(out["final_boards"]), want["final_boards"])
                        and np.array_equal(out["lengths"].cpu().numpy(), want["lengths"])
                        and np.array_equal(out["scores"].cpu().numpy(), want["scores"]))
                steps = int(want["lengths"].sum())
                total_steps += steps
                bad += 0 if same else 1
                print(f"mode {mode} policy {policy} seed {s}: {steps} env-steps, {'identical' if same else 'MISMATCH'}", flush=True)
    print(f"soak: {total_steps} env-steps in {args.seeds * 4} batches of {args.envs} envs, {bad} mismatching batches, "
          f"{time.perf_counter() - t0:.0f} s")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
