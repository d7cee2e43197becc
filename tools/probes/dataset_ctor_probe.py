"""Where does the construction of DevicePPOBatches (C4: 3.1e7 kept steps) spend its time?  Usage: python tools/probes/dataset_ctor_probe.py"""
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path[:0] = [str(ROOT), str(ROOT / "2048-ppo-agent_b200")]

import torch

import g2048
from g2048 import engine as E
from g2048.ppo import data_loader as DL


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best * 1e3


def main():
    dev = torch.device("cuda:0")
    n = 31_000_000
    packed = dict(boards=torch.randint(0, 1 << 62, (n,), dtype=torch.int64, device=dev),
                  meta=((torch.rand(n, device=dev) < 1 / 118).to(torch.uint8) << 6),
                  rewards=torch.rand(n, device=dev), log_probs=torch.rand(n, device=dev), values=torch.rand(n, device=dev))
    dones = DL.meta_to_dones(packed["meta"])
    print(f"meta_to_dones        {timed(lambda: DL.meta_to_dones(packed['meta'])):8.3f} ms")
    print(f"compute_gae (+norm)  {timed(lambda: DL.compute_gae(packed['rewards'], packed['values'], dones)):8.3f} ms")
    print(f"  gae_flat only      {timed(lambda: E.gae_flat(packed['rewards'], packed['values'], dones, 0.99, 0.95)):8.3f} ms")
    print(f"randperm(n)[:300000] {timed(lambda: torch.randperm(n, device=dev)[:300000]):8.3f} ms")
    print(f"whole constructor    {timed(lambda: g2048.DevicePPOBatches(packed, 0.99, 0.95, batch_size=2048, max_samples_per_epoch=300000, shuffle_on_reset=True)):8.3f} ms")
    if hasattr(E, "random_subset"):
        print(f"random_subset        {timed(lambda: E.random_subset(n, 300000, 7, dev)):8.3f} ms")


if __name__ == "__main__":
    main()
