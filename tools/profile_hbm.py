"""One launch of every HBM-bound kernel at the sizes bench.py's `roofline_hbm` rows use, for an `ncu --set full` capture:

    ncu --set full --clock-control none --import-source on -k regex:'expand_obs|gae_flat4_kernel|normalize|gae_time_major|embed_boards_kernel|embed_grad_partial' \\
        -o gpurun_out/hbm_rows python tools/profile_hbm.py
    python tools/summarize_ncu.py gpurun_out/hbm_rows.ncu-rep profiles/r01o_hbm_rows

The launch order is the order of the rows; profiles/hbm_traffic.json (dram bytes read + written per launch, taken
from that capture) is what bench.py reports as `traffic` for them."""
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

    def cold():
        flush.fill_(3)  # a torch kernel: not captured by the -k filter
        torch.cuda.synchronize()

    n_b = 1 << 20
    boards = torch.randint(0, 1 << 62, (n_b,), dtype=torch.int64, device=dev)
    out = torch.empty((n_b, 16, 31), dtype=torch.float32, device=dev)
    cold()
    N.call("g2048_expand_obs", N.ptr(boards), n_b, N.OBS_F32, N.ptr(out), 0, 0, N.stream_ptr())  # 1: expand_obs f32
    del out

    n_buf, m = 1 << 22, 1 << 16
    packed = dict(boards=torch.randint(0, 1 << 62, (n_buf,), dtype=torch.int64, device=dev),
                  meta=torch.randint(0, 127, (n_buf,), dtype=torch.uint8, device=dev),
                  log_probs=torch.rand(n_buf, device=dev), values=torch.rand(n_buf, device=dev))
    g_adv, g_ret = torch.rand(n_buf, device=dev), torch.rand(n_buf, device=dev)
    for mm in (m, 1 << 19):
        idx = torch.randint(0, n_buf, (mm,), device=dev)
        mb = E.minibatch_buffers(mm, dev)
        cold()
        E.gather_minibatch(idx, packed, g_adv, g_ret, out=mb)  # 2, 3: gathered expand_obs with scalars
        del mb, idx
    del packed, g_adv, g_ret

    n_e, d_model = 1 << 18, 256
    eb = boards[:n_e].contiguous()
    for dt in (torch.float32, torch.bfloat16):
        table = torch.randn(31, d_model, device=dev).to(dt)
        emb = torch.empty((n_e, 16, d_model), dtype=dt, device=dev)
        cold()
        E.embed_boards(eb, table, out=emb)  # 4, 6: embed_boards_kernel
        cold()
        E.embed_boards_grad(eb, emb)  # 5, 7: embed_grad_partial_kernel
        del emb
    del boards, eb

    n_g = 1 << 26
    r, v = torch.rand(n_g, device=dev), torch.rand(n_g, device=dev)
    d = (torch.rand(n_g, device=dev) < 1 / 300).to(torch.uint8)
    adv, ret = torch.empty(n_g, device=dev), torch.empty(n_g, device=dev)
    scratch = torch.zeros(int(N.lib.g2048_gae_flat_scratch_bytes(n_g)), dtype=torch.uint8, device=dev)
    mom = torch.zeros(6, dtype=torch.float64, device=dev)
    cold()
    N.call("g2048_gae_flat", N.ptr(r), N.ptr(v), N.ptr(d), n_g, 0.99, 0.95, N.ptr(adv), N.ptr(ret), N.ptr(scratch),
           N.ptr(mom), N.stream_ptr())  # 8: gae_flat4_kernel, constant done rate
    cold()
    N.call("g2048_normalize", N.ptr(adv), n_g, N.ptr(mom), 1, N.stream_ptr())  # 9: normalize_kernel
    subs = E.chain_advance(E.words_tensor([0, 2048], dev), E.RNG_PARTITIONABLE, 1 + 2 * 2048)
    lens = E.play(N.POLICY_DRUL, subs, 1 << 18, 0, 1 << 18, E.RNG_PARTITIONABLE)["lengths"].to(torch.int64)
    ends = torch.cumsum(lens, 0) - 1
    period = int(ends[-1]) + 1
    reps = (n_g + period - 1) // period
    d_real = torch.zeros(reps * period, dtype=torch.uint8, device=dev)
    d_real.view(reps, period)[:, ends] = 1
    d_real = d_real[:n_g].contiguous()
    scratch.zero_()
    cold()
    N.call("g2048_gae_flat", N.ptr(r), N.ptr(v), N.ptr(d_real), n_g, 0.99, 0.95, N.ptr(adv), N.ptr(ret), N.ptr(scratch),
           N.ptr(mom), N.stream_ptr())  # 10: gae_flat4_kernel, real episode lengths
    del r, v, d, d_real, adv, ret

    for b in (1 << 16, 1 << 18):
        t_steps = 128
        rr, vv = torch.rand((t_steps, b), device=dev), torch.rand((t_steps, b), device=dev)
        mm = ((torch.rand((t_steps, b), device=dev) < 1 / 300).to(torch.uint8) << 6)
        a2, r2 = torch.empty((t_steps, b), device=dev), torch.empty((t_steps, b), device=dev)
        cold()
        N.call("g2048_gae_time_major", N.ptr(rr), N.ptr(vv), N.ptr(mm), t_steps, b, None, 0.99, 0.95, N.ptr(a2), N.ptr(r2),
               N.ptr(mom), N.stream_ptr())  # 11, 12: gae_time_major_kernel
        del rr, vv, mm, a2, r2
    torch.cuda.synchronize()
    print("done")


if __name__ == "__main__":
    main()
