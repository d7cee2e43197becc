"""The reference-API calls at C1's size (1 024 envs) for a launch list:
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/probes/api_launches.py"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "2048-ppo-agent_b200")]
import torch

import g2048
from g2048.runs.run_actions_max_tile import run_actions_max_tile

for rep in range(2):
    torch.cuda.nvtx.range_push(f"rep{rep}")
    stats = run_actions_max_tile(0, 1024, 4096, g2048.act_randomly)
    states = g2048.BatchRunner(init_seed=0, act_fn=g2048.act_randomly).run_rollout_batch(1024)
    out = g2048.BatchRunner(init_seed=0, act_fn=g2048.act_drul, pinned_outputs=True).run_actions_batch(1024)
    summary = g2048.BatchRunner(init_seed=1, act_fn=g2048.act_drul).run_stats_batch(1024)["summary"]
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_pop()
print(stats.mean, len(states), out[0].shape, summary["env_steps"])
