"""One PPO iteration with a small network policy (4 096 envs to termination, 2 epochs of 2 048-sample minibatches) for a
launch list: ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/probes/net_iteration_launches.py"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "2048-ppo-agent_b200")]
import torch

import g2048
from g2048.ppo import PPOIterationLoop


class Agent(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.body = torch.nn.Linear(496, 64)
        self.head = torch.nn.Linear(64, 5)

    def forward(self, obs, mask=None):
        out = self.head(torch.tanh(self.body(obs.reshape(obs.shape[0], -1).float())))
        return out[:, :4], out[:, 4:5]


torch.manual_seed(0)
agent = Agent().cuda()
act = g2048.TorchActionFunction(agent, use_mask=True, device=torch.device("cuda"))
runner = g2048.BatchRunner(init_seed=0, act_fn=act)
loop = PPOIterationLoop(runner, g2048.RolloutBuffer(31, 16, 4), minibatch_step=lambda b: {"kl": 0.0}, agent=agent)
for it in range(2):
    ro = loop.collect_rollouts(4096, 1)
    up = loop.update_policy(batch_size=2048, n_epochs=2)
    torch.cuda.synchronize()
    print(it, ro, up)
