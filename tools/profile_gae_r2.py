"""One launch each of the round-2 GAE kernels for an ncu capture: g2048_gae_flat_scan at 2^26 steps and
g2048_gae_time_major (T split over the warps of a CTA) at C3 (128 x 65536)."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "2048-ppo-agent_b200")]
import torch

from g2048 import _native as N


def main():
    dev = torch.device("cuda:0")
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    n_g = 1 << 26
    r, v = torch.rand(n_g, device=dev), torch.rand(n_g, device=dev)
    d = (torch.rand(n_g, device=dev) < 1 / 300).to(torch.uint8)
    adv, ret = torch.empty(n_g, device=dev), torch.empty(n_g, device=dev)
    scratch = torch.zeros(int(N.lib.g2048_gae_scan_scratch_bytes(n_g)), dtype=torch.uint8, device=dev)
    mom = torch.zeros(6, dtype=torch.float64, device=dev)
    flush.fill_(1)
    torch.cuda.synchronize()
    N.call("g2048_gae_flat_scan", N.ptr(r), N.ptr(v), N.ptr(d), n_g, 0.99, 0.95, N.ptr(adv), N.ptr(ret), N.ptr(scratch), N.ptr(mom),
           N.stream_ptr())
    del r, v, d, adv, ret
    t_steps, b = 128, 1 << 16
    rr, vv = torch.rand((t_steps, b), device=dev), torch.rand((t_steps, b), device=dev)
    mm = ((torch.rand((t_steps, b), device=dev) < 1 / 300).to(torch.uint8) << 6)
    a2, r2 = torch.empty((t_steps, b), device=dev), torch.empty((t_steps, b), device=dev)
    flush.fill_(1)
    torch.cuda.synchronize()
    N.call("g2048_gae_time_major", N.ptr(rr), N.ptr(vv), N.ptr(mm), t_steps, b, None, 0.99, 0.95, N.ptr(a2), N.ptr(r2), N.ptr(mom),
           N.stream_ptr())
    torch.cuda.synchronize()
    print("done")


if __name__ == "__main__":
    main()
