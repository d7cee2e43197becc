"""Eager vs CUDA-graph vs live-env-compacting rollout loop with a network policy.
A small transformer stands in for the reference's PPOAgent (d_model 256, 4 layers); the network is PyTorch either
way -- what changes is launch overhead and the per-step host synchronisation.  Prints one JSON object."""
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "2048-ppo-agent_b200")]
import torch

import g2048


class Agent(torch.nn.Module):
    def __init__(self, d_model=256, layers=4):
        super().__init__()
        self.input_embedding = torch.nn.Linear(31, d_model, bias=False)
        layer = torch.nn.TransformerEncoderLayer(d_model, 8, 1024, dropout=0.0, batch_first=True)
        self.encoder = torch.nn.TransformerEncoder(layer, layers, enable_nested_tensor=False)
        self.actor = torch.nn.Linear(d_model, 4, bias=False)
        self.critic = torch.nn.Linear(d_model, 1, bias=False)

    def forward(self, obs, mask=None):
        f = self.encoder(self.input_embedding(obs)).mean(dim=1)
        return self.actor(f), self.critic(f)


class SmallAgent(torch.nn.Module):
    """Launch-bound case: three small GEMMs per step."""

    def __init__(self):
        super().__init__()
        self.body = torch.nn.Sequential(torch.nn.Flatten(), torch.nn.Linear(496, 256), torch.nn.ReLU())
        self.actor = torch.nn.Linear(256, 4, bias=False)
        self.critic = torch.nn.Linear(256, 1, bias=False)

    def forward(self, obs, mask=None):
        h = self.body(obs)
        return self.actor(h), self.critic(h)


def main():
    out = {}
    for make, n in ((SmallAgent, 1024), (SmallAgent, 16384), (Agent, 256), (Agent, 4096)):
        row = {}
        for mode in ("eager", "graph", "compact"):
            graph = mode == "graph"
            torch.manual_seed(0)
            fn = g2048.TorchActionFunction(make().cuda(), use_mask=True, device=torch.device("cuda"))
            runner = g2048.BatchRunner(init_seed=2, act_fn=fn, cuda_graph=graph, compact_live=mode == "compact")
            runner.run_packed_batch(n)  # warm-up (+ capture)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            ro = runner.run_packed_batch(n)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            row[mode] = {"seconds": dt, "loop_steps": ro.t_steps, "ms_per_loop_step": 1e3 * dt / ro.t_steps,
                         "env_steps_per_sec": ro.env_steps / dt, "live_fraction": ro.env_steps / (ro.t_steps * n)}
        row["speedup"] = row["eager"]["seconds"] / row["graph"]["seconds"] * row["graph"]["loop_steps"] / row["eager"]["loop_steps"]
        row["speedup_compact"] = row["eager"]["ms_per_loop_step"] / row["compact"]["ms_per_loop_step"]
        out[f"{make.__name__},envs={n}"] = row
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
