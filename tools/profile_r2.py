"""One launch of each kernel that is new in round 2, at the sizes bench.py measures them, for an `ncu --set full` capture:

    python tools/profile_r2.py && ncu --set full --clock-control none --import-source on \
        -k regex:'gae_scan_kernel|play_record_compact|expand_obs_tma_kernel|pack_samples|policy_step_obs|play3_kernel' \
        -o gpurun_out/r02_kernels python tools/profile_r2.py
    python tools/summarize_ncu.py gpurun_out/r02_kernels.ncu-rep profiles/r02_new_kernels

Launch order = row order of the summary."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "2048-ppo-agent_b200")]

import torch

from g2048 import _native as N
from g2048 import engine as E


def main():
    dev = torch.device("cuda:0")
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    mode = E.RNG_PARTITIONABLE

    def cold():
        flush.fill_(3)  # a torch kernel: not captured by the -k filter
        torch.cuda.synchronize()

    # 1: recording play kernel, C4 size; 2: its compaction into the flat buffer
    subs = E.chain_advance(E.words_tensor([0, 4], dev), mode, 1 + 2 * 2048)
    n = 1 << 18
    cold()
    rec = E.play_record(E.POLICY_RANDOM, subs, n, 0, n, mode)
    offsets = E.exclusive_scan(rec["lengths"])
    total = int(offsets[-1])
    cold()
    flat = E.play_record_compact(rec, offsets, total)
    del rec

    # 3: GAE as a reverse scan, 2^26 steps
    n_g = 1 << 26
    r, v = torch.rand(n_g, device=dev), torch.rand(n_g, device=dev)
    d = (torch.rand(n_g, device=dev) < 1 / 300).to(torch.uint8)
    adv, ret = torch.empty(n_g, device=dev), torch.empty(n_g, device=dev)
    scratch = torch.zeros(int(N.lib.g2048_gae_scan_scratch_bytes(n_g)), dtype=torch.uint8, device=dev)
    mom = torch.zeros(6, dtype=torch.float64, device=dev)
    cold()
    N.call("g2048_gae_flat_scan", N.ptr(r), N.ptr(v), N.ptr(d), n_g, 0.99, 0.95, N.ptr(adv), N.ptr(ret), N.ptr(scratch),
           N.ptr(mom), N.stream_ptr())
    del r, v, d, adv, ret

    # 4: sample records; 5, 6: minibatch gather from them (65 536 and 2^19 samples)
    n_buf = 1 << 22
    packed = dict(boards=torch.randint(0, 1 << 62, (n_buf,), dtype=torch.int64, device=dev),
                  meta=torch.randint(0, 127, (n_buf,), dtype=torch.uint8, device=dev), rewards=torch.rand(n_buf, device=dev),
                  log_probs=torch.rand(n_buf, device=dev), values=torch.rand(n_buf, device=dev))
    g_adv, g_ret = torch.rand(n_buf, device=dev), torch.rand(n_buf, device=dev)
    cold()
    records = E.pack_samples(packed, g_adv, g_ret, None)
    for m in (1 << 16, 1 << 19):
        idx = torch.randint(0, n_buf, (m,), device=dev)
        out = E.minibatch_buffers(m, dev)
        cold()
        E.gather_samples(idx, records, out=out)
        del out, idx
    del packed, records, g_adv, g_ret

    # 7: fused policy step + next observation, C3 size
    b = 1 << 16
    subs3 = E.chain_advance(E.words_tensor([0, 3], dev), mode, 5)
    pb, ps = E.env_init(subs3[0], b, 0, b, mode)
    logits, values = torch.randn((b, 4), device=dev), torch.randn(b, device=dev)
    obs = torch.empty((b, 16, 31), dtype=torch.float32, device=dev)
    rec_b = torch.empty(b, dtype=torch.int64, device=dev)
    rec_m = torch.empty(b, dtype=torch.uint8, device=dev)
    rec_r, rec_l, rec_v = (torch.empty(b, dtype=torch.float32, device=dev) for _ in range(3))
    cold()
    E.policy_step_obs(pb, ps, logits, values, True, True, True, subs3[1:], None, b, 0, mode, obs, rec_b, rec_m, rec_r, rec_l, rec_v)
    torch.cuda.synchronize()
    print("done", total, float(flat["rewards"].sum()))


if __name__ == "__main__":
    main()
