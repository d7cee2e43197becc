"""BASELINE.json config C4 -- the data path of one full PPO iteration at 262 144 envs, phase by phase, on one GPU: the same
leg `bench.py` reports as `c4_iteration` (recorded rollout + rollout-buffer write through `run_flat_batch` / `store_flat`,
round 1's lock-step recorder + `store_packed` beside it, GAE + sample records, four epochs of minibatches, the bit-exact
board check on a 4 096-env sample against the oracle).  Prints one JSON object.  Usage: python tools/c4_iteration.py"""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "2048-ppo-agent_b200")]

import torch

import bench
from g2048 import _native as N
from g2048 import engine as E

if __name__ == "__main__":
    torch.cuda.set_device(0)
    report = bench.c4_iteration(E, N, torch, torch.device("cuda", 0), bench.measured_hbm_peak()[0])
    print(json.dumps(report))
    assert report["bit_exact_board_check"]["identical"], "recorded rollout differs from the oracle"
