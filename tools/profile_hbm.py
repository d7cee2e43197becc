"""One launch of the dominant kernel of every row of bench.py's `roofline_hbm` (plus the C4 compaction), at the bench's
sizes, L2 flushed before each -- the target of an `ncu --set full` capture:

    python tools/profile_hbm.py run &&
    ncu --set full --clock-control none --import-source on -k regex:"$(python tools/profile_hbm.py regex)" \
        -o gpurun_out/r02_hbm_rows python tools/profile_hbm.py run
    python tools/summarize_ncu.py gpurun_out/r02_hbm_rows.ncu-rep profiles/r02_hbm_rows      # here
    python tools/profile_hbm.py parse gpurun_out/r02_hbm_rows.ncu-rep                         # -> profiles/hbm_traffic.json

ROWS is the launch order: (bench row name, substring of the kernel that the row is about).  profiles/hbm_traffic.json
(dram bytes read + written per launch) is what bench.py reports as `traffic` next to the algorithmic bytes.
"""
import csv
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "2048-ppo-agent_b200")]

ROWS = [
    ("expand_obs_tma_kernel<float>", "expand_obs_tma_kernel"),
    ("pack_samples_kernel", "pack_samples_kernel"),
    ("gather_samples (gather_samples_tile_kernel<float>)", "gather_samples_tile_kernel"),
    ("gather_minibatch (round 1: one source array per field)", "expand_obs_tma_kernel"),
    ("gather_samples, 2^19 samples", "gather_samples_tile_kernel"),
    ("embed_boards (float32)", "embed_boards_kernel"),
    ("embed_grad_partial_kernel + reduce (float32)", "embed_grad_partial_kernel"),
    ("embed_boards (bfloat16)", "embed_boards_kernel"),
    ("embed_grad_partial_kernel + reduce (bfloat16)", "embed_grad_partial_kernel"),
    ("gae_flat4_kernel (pipelined)", "gae_flat4_kernel"),
    ("gae_scan_kernel (opt-in: re-associated reverse scan, 1e-5 relative)", "gae_scan_kernel"),
    ("normalize_kernel", "normalize_kernel"),
    ("gae_flat4_kernel, episode lengths of real play", "gae_flat4_kernel"),
    ("gae_time_major_kernel", "gae_time_major_"),
    ("gae_time_major_kernel, 4 x C3", "gae_time_major_"),
    ("c4: play_record_compact_kernel", "play_record_compact_kernel"),
    ("c3: policy_step_obs_kernel<float>", "policy_step_obs_kernel"),
]
REGEX = "expand_obs_tma_kernel|gather_samples_tile_kernel|pack_samples_kernel|embed_boards_kernel|embed_grad_partial_kernel|gae_flat4_kernel|gae_scan_kernel|" \
        "normalize_kernel|gae_time_major_|play_record_compact_kernel|policy_step_obs_kernel"


def run():
    import torch

    from g2048 import _native as N
    from g2048 import engine as E

    dev = torch.device("cuda:0")
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

    def cold():
        flush.fill_(3)  # a torch kernel: not captured by the -k filter
        torch.cuda.synchronize()

    n_b = 1 << 20
    boards = torch.randint(0, 1 << 62, (n_b,), dtype=torch.int64, device=dev)
    out = torch.empty((n_b, 16, 31), dtype=torch.float32, device=dev)
    cold()
    N.call("g2048_expand_obs", N.ptr(boards), n_b, N.OBS_F32, N.ptr(out), 0, 0, N.stream_ptr())
    del out

    n_buf, m = 1 << 22, 1 << 16
    packed = dict(boards=torch.randint(0, 1 << 62, (n_buf,), dtype=torch.int64, device=dev),
                  meta=torch.randint(0, 127, (n_buf,), dtype=torch.uint8, device=dev), rewards=torch.rand(n_buf, device=dev),
                  log_probs=torch.rand(n_buf, device=dev), values=torch.rand(n_buf, device=dev))
    g_adv, g_ret = torch.rand(n_buf, device=dev), torch.rand(n_buf, device=dev)
    mom_p = torch.tensor([n_buf, 0.5 * n_buf, 0.34 * n_buf, 0.5 * n_buf, 0.34 * n_buf, 0.0], dtype=torch.float64, device=dev)
    cold()
    records = E.pack_samples(packed, g_adv, g_ret, mom_p)
    idx = torch.randperm(n_buf, device=dev)[:m].contiguous()
    mb = E.minibatch_buffers(m, dev)
    cold()
    E.gather_samples(idx, records, out=mb)
    cold()
    E.gather_minibatch(idx, packed, g_adv, g_ret, out=mb)
    del mb, idx
    idx = torch.randint(0, n_buf, (1 << 19,), device=dev)
    mb = E.minibatch_buffers(1 << 19, dev)
    cold()
    E.gather_samples(idx, records, out=mb)
    del mb, idx, packed, records, g_adv, g_ret

    n_e, d_model = 1 << 18, 256
    eb = boards[:n_e].contiguous()
    for dt in (torch.float32, torch.bfloat16):
        table = torch.randn(31, d_model, device=dev).to(dt)
        emb = torch.empty((n_e, 16, d_model), dtype=dt, device=dev)
        cold()
        E.embed_boards(eb, table, out=emb)
        cold()
        E.embed_boards_grad(eb, emb)
        del emb
    del boards, eb

    n_g = 1 << 26
    r, v = torch.rand(n_g, device=dev), torch.rand(n_g, device=dev)
    d = (torch.rand(n_g, device=dev) < 1 / 300).to(torch.uint8)
    adv, ret = torch.empty(n_g, device=dev), torch.empty(n_g, device=dev)
    scratch = torch.zeros(int(N.lib.g2048_gae_flat_scratch_bytes(n_g)), dtype=torch.uint8, device=dev)
    scan_scratch = torch.zeros(int(N.lib.g2048_gae_scan_scratch_bytes(n_g)), dtype=torch.uint8, device=dev)
    mom = torch.zeros(6, dtype=torch.float64, device=dev)
    cold()
    N.call("g2048_gae_flat", N.ptr(r), N.ptr(v), N.ptr(d), n_g, 0.99, 0.95, N.ptr(adv), N.ptr(ret), N.ptr(scratch), N.ptr(mom),
           N.stream_ptr())
    cold()
    N.call("g2048_gae_flat_scan", N.ptr(r), N.ptr(v), N.ptr(d), n_g, 0.99, 0.95, N.ptr(adv), N.ptr(ret), N.ptr(scan_scratch), N.ptr(mom),
           N.stream_ptr())
    cold()
    N.call("g2048_normalize", N.ptr(adv), n_g, N.ptr(mom), 1, N.stream_ptr())
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
    N.call("g2048_gae_flat", N.ptr(r), N.ptr(v), N.ptr(d_real), n_g, 0.99, 0.95, N.ptr(adv), N.ptr(ret), N.ptr(scratch), N.ptr(mom),
           N.stream_ptr())
    del r, v, d, d_real, adv, ret

    for b in (1 << 16, 1 << 18):
        t_steps = 128
        rr, vv = torch.rand((t_steps, b), device=dev), torch.rand((t_steps, b), device=dev)
        mm = ((torch.rand((t_steps, b), device=dev) < 1 / 300).to(torch.uint8) << 6)
        a2, r2 = torch.empty((t_steps, b), device=dev), torch.empty((t_steps, b), device=dev)
        cold()
        N.call("g2048_gae_time_major", N.ptr(rr), N.ptr(vv), N.ptr(mm), t_steps, b, None, 0.99, 0.95, N.ptr(a2), N.ptr(r2), N.ptr(mom),
               N.stream_ptr())
        del rr, vv, mm, a2, r2

    # C4: the compaction of a recorded batch (the recording kernel itself is a play3_kernel: see tools/inst_counts.py)
    n = 1 << 18
    rec = E.play_record(E.POLICY_RANDOM, subs, n, 0, n, E.RNG_PARTITIONABLE)
    offsets = E.exclusive_scan(rec["lengths"])
    total = int(offsets[-1])
    cold()
    E.play_record_compact(rec, offsets, total)
    del rec
    # C3: the fused policy step + next observation
    b = 1 << 16
    pb, ps = E.env_init(subs[0], b, 0, b, E.RNG_PARTITIONABLE)
    logits, values = torch.randn((b, 4), device=dev), torch.randn(b, device=dev)
    obs = torch.empty((b, 16, 31), dtype=torch.float32, device=dev)
    rec_b = torch.empty(b, dtype=torch.int64, device=dev)
    rec_m = torch.empty(b, dtype=torch.uint8, device=dev)
    rec_r, rec_l, rec_v = (torch.empty(b, dtype=torch.float32, device=dev) for _ in range(3))
    cold()
    E.policy_step_obs(pb, ps, logits, values, True, True, True, subs[1:], None, b, 0, E.RNG_PARTITIONABLE, obs, rec_b, rec_m, rec_r,
                      rec_l, rec_v)
    torch.cuda.synchronize()
    print("done", total)


def parse(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {name: hdr.index(name) for name in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum")}
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "nsecond": 1e-3, "usecond": 1, "msecond": 1e3}

    def val(r, name):
        return float(r[col[name]].replace(",", "")) * scale.get(units[col[name]], 1)

    out, k = {}, 0
    for r in data:  # launches in order; a row takes the next launch whose kernel name matches (others are helper kernels)
        if k >= len(ROWS):
            break
        name, needle = ROWS[k]
        if needle not in r[col["Kernel Name"]]:
            continue
        rd, wr = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
        out[name] = {"traffic": rd + wr, "dram_read": rd, "dram_write": wr, "kernel": r[col["Kernel Name"]][:60],
                     "ncu_duration_us": val(r, "gpu__time_duration.sum")}
        k += 1
    assert k == len(ROWS), f"only {k} of {len(ROWS)} rows found in {rep}"
    doc = {"source": f"{Path(rep).name} (ncu --set full --clock-control none on tools/profile_hbm.py: one launch per row at bench.py's "
                     "sizes, L2 flushed before each); summary in profiles/r02_hbm_rows.csv",
           "note": "dram__bytes_read.sum + dram__bytes_write.sum of the row's dominant kernel, bytes per launch; writes still in the "
                   "126 MB L2 when the kernel ends are not counted by ncu, so write-heavy rows can read below their algorithmic bytes",
           "rows": out}
    (ROOT / "profiles" / "hbm_traffic.json").write_text(json.dumps(doc, indent=1) + "\n")
    for name, v in out.items():
        print(f"{name[:60]:60s} read {v['dram_read'] / 1e6:9.1f} MB  write {v['dram_write'] / 1e6:9.1f} MB  {v['ncu_duration_us']:8.1f} us")


if __name__ == "__main__":
    cmd = sys.argv[1] if len(sys.argv) > 1 else "run"
    if cmd == "run":
        run()
    elif cmd == "regex":
        print(REGEX)
    else:
        parse(sys.argv[2])
