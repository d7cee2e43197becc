"""Flat GAE timing at 2^26 steps for one or more builds of libg2048.so (development aid).

    python tools/probes/gae_ab.py [path/to/libg2048_variant.so ...]

Each library is loaded with ctypes on its own and times g2048_gae_flat on the same three buffers: dones at a
constant rate of 1/300 (geometric lengths), the episode lengths of real DRUL games, and a rate of 1/1000; L2 is
flushed between launches.  The first library's outputs are the reference the others must equal bit for bit."""
import ctypes as C
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path[:0] = [str(ROOT), str(ROOT / "2048-ppo-agent_b200")]

import torch

from g2048 import engine as E
from g2048 import _native as N


def main():
    libs = [str(N.LIB_PATH)] + sys.argv[1:]
    dev = torch.device("cuda:0")
    n = 1 << 26
    torch.manual_seed(0)
    r, v = torch.rand(n, device=dev), torch.rand(n, device=dev)
    cases = {"rate 1/300": (torch.rand(n, device=dev) < 1 / 300).to(torch.uint8),
             "rate 1/1000": (torch.rand(n, device=dev) < 1 / 1000).to(torch.uint8)}
    subs = E.chain_advance(E.words_tensor([0, 2048], dev), E.RNG_PARTITIONABLE, 1 + 2 * 2048)
    lens = E.play(N.POLICY_DRUL, subs, 1 << 18, 0, 1 << 18, E.RNG_PARTITIONABLE)["lengths"].to(torch.int64)
    ends = torch.cumsum(lens, 0) - 1
    period = int(ends[-1]) + 1
    reps = (n + period - 1) // period
    d_real = torch.zeros(reps * period, dtype=torch.uint8, device=dev)
    d_real.view(reps, period)[:, ends] = 1
    cases["real DRUL lengths"] = d_real[:n].contiguous()
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    want = {}
    for path in libs:
        lib = C.CDLL(path)
        lib.g2048_gae_flat.restype = C.c_int
        lib.g2048_gae_flat.argtypes = [C.c_void_p] * 3 + [C.c_int64, C.c_double, C.c_double] + [C.c_void_p] * 5
        lib.g2048_gae_flat_scratch_bytes.restype = C.c_int64
        lib.g2048_gae_flat_scratch_bytes.argtypes = [C.c_int64]
        scratch = torch.zeros(int(lib.g2048_gae_flat_scratch_bytes(n)), dtype=torch.uint8, device=dev)
        adv, ret = torch.empty(n, device=dev), torch.empty(n, device=dev)
        mom = torch.zeros(6, dtype=torch.float64, device=dev)
        line = [Path(path).name]
        for name, d in cases.items():
            def run():
                scratch.zero_()
                rc = lib.g2048_gae_flat(r.data_ptr(), v.data_ptr(), d.data_ptr(), n, 0.99, 0.95, adv.data_ptr(), ret.data_ptr(),
                                        scratch.data_ptr(), mom.data_ptr(), 0)
                assert rc == 0, rc
            run()
            torch.cuda.synchronize()
            ts = []
            for _ in range(7):
                flush.fill_(1)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); run(); b.record(); torch.cuda.synchronize()
                ts.append(a.elapsed_time(b) * 1e3)
            ts.sort()
            if name not in want:
                want[name] = (adv.clone(), ret.clone())
                same = "ref"
            else:
                same = "same" if torch.equal(adv, want[name][0]) and torch.equal(ret, want[name][1]) else "DIFFERENT"
            line.append(f"{name}: {ts[len(ts) // 2]:.1f} us ({n * 17 / ts[len(ts) // 2] / 1e6:.2f} TB/s, {same})")
        print(" | ".join(line), flush=True)


if __name__ == "__main__":
    main()
