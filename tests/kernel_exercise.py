"""Every kernel and C entry once at small, ragged sizes, with every output tensor allocated through `T`
(torch itself, or tests/test_guard_bands_gpu.py's guarded allocator that surrounds each output with canary bytes).

compute-sanitizer is not available on the GPU pool, so out-of-bounds writes are looked for this way instead:
odd sizes that do not divide any tile / vector width, canaries either side of every output, and (elsewhere in
tests/) bit-exact comparison of the outputs with the oracle.
"""
import numpy as np
import torch


def run_all(E, g2048, T=torch, dev="cuda"):
    for mode in (0, 1):
        key = E.words_tensor([0, 5], dev)
        subs = E.chain_advance(key, mode, 1 + 2 * 700)
        E.split_keys(subs[0], 1000, 17, 333, mode)
        E.threefry2x32(subs[:77].contiguous(), subs[100:177].contiguous())
        boards, status = E.env_init(subs[0], 1000, 17, 333, mode)
        a, _ = E.act(0, status, subs[1], 1000, 17, mode)
        E.act(1, status, subs[1], 1000, 17, mode)
        E.env_step(boards.clone(), status.clone(), a, subs[2], 1000, 17, mode)
        E.env_step_draws(boards.clone(), status.clone(), a, torch.zeros(333, dtype=torch.int32, device=dev),
                         torch.ones(333, dtype=torch.int32, device=dev))
        for entry in ("g2048_play_swar", "g2048_play_v1"):
            for policy in (0, 1):
                E.play(policy, subs, 777, 100, 300, mode, entry=entry)
        for policy in (0, 1):
            E.play(policy, subs, 3001, 0, 3001, mode, entry="g2048_play_tables")
            E.play(policy, subs, 3001, 1234, 1, mode, entry="g2048_play_tables")
        for policy in (0, 1):  # recording form: arena + compaction into the flat buffer
            rec = E.play_record(policy, subs, 3001, 100, 777, mode)
            offs_r = E.exclusive_scan(rec["lengths"])
            E.play_record_compact(rec, offs_r, int(offs_r[-1].item()))
        counters = T.zeros(4, dtype=torch.int64, device=dev)
        n, ch = 333, 5
        rb = T.empty((ch, n), dtype=torch.int64, device=dev)
        rm = T.empty((ch, n), dtype=torch.uint8, device=dev)
        rr = T.empty((ch, n), dtype=torch.float32, device=dev)
        rl = T.empty((ch, n), dtype=torch.float32, device=dev)
        E.rollout_steps(0, boards, status, subs[3:], ch, 0, 1000, 17, mode, rb, rm, rr, rl, counters)
        live_all = torch.nonzero((status & 16) == 0).flatten()[1::2].contiguous()
        if live_all.numel():
            E.rollout_steps_live(1, boards, status, subs[3:], ch, 0, 1000, 17, mode, live_all, rb, rm, rr, None, counters)
        logits = torch.randn(n, 4, device=dev)
        values = torch.randn(n, device=dev)
        rv = T.empty(n, dtype=torch.float32, device=dev)
        acts = T.empty(n, dtype=torch.int32, device=dev)
        E.policy_step(boards, status, logits, values, True, True, True, subs[5], subs[6], 1000, 17, mode,
                      rb[0], rm[0], rr[0], rl[0], rv, acts)
        step_index = T.zeros((), dtype=torch.int32, device=dev)
        E.counter_add(step_index, ch - 1)  # last record slot
        E.policy_step_at(boards, status, logits, values, True, True, True, subs[7:].contiguous(), step_index, 1000, 17, mode,
                         rb, rm, rr, rl, T.empty((ch, n), dtype=torch.float32, device=dev))
        for dt in (torch.float32, torch.bfloat16, torch.bool):  # fused step + next observation
            obs_next = T.empty((n, 16, 31), dtype=dt, device=dev)
            E.policy_step_obs(boards, status, logits, values, True, True, True, subs[7:].contiguous(), None, 1000, 17, mode,
                              obs_next, rb[0], rm[0], rr[0], rl[0], rv, acts, T.zeros(2, dtype=torch.int64, device=dev))
        live = torch.nonzero((status & 16) == 0).flatten()[::3].contiguous()
        if live.numel():
            E.expand_obs_gather(boards, live, torch.float32)
            E.expand_obs_gather(boards, live, torch.bool)
            E.policy_step_live(boards, status, torch.randn(live.numel(), 4, device=dev), torch.randn(live.numel(), device=dev),
                               True, True, False, subs[5], subs[6], live, 1000, 17, mode, rb[1], rm[1], rr[1], rl[1],
                               T.zeros((n,), dtype=torch.float32, device=dev))
        E.sample_logits(logits, status, True, True, subs[5], 1000, 17, mode, want_entropy=True)
        E.evaluate_logits(logits, status, True, acts)
        E.unpack_records(rm, rr, rl, None, ch, n)
        lengths = E.episode_lengths(rm, ch, n)
        offs = E.exclusive_scan(lengths)
        tot = int(offs[-1].item())
        if tot:
            fb = T.empty(tot, dtype=torch.int64, device=dev)
            fm = T.empty(tot, dtype=torch.uint8, device=dev)
            fr = T.empty(tot, dtype=torch.float32, device=dev)
            fl = T.empty(tot, dtype=torch.float32, device=dev)
            fv = T.zeros(tot, dtype=torch.float32, device=dev)
            E.compact_records(rb, rm, rr, rl, None, ch, n, lengths, offs, 0, fb, fm, fr, fl, fv)
            E.unpack_flat_meta(fm)
            E.meta_dones(fm)
        for dt in (torch.float32, torch.bfloat16, torch.bool):
            for entry in ("g2048_expand_obs", "g2048_expand_obs_v1"):
                E.expand_obs(boards, dt, entry=entry)
                E.expand_obs(rb, dt, rows=ch, n_cols=n, entry=entry)
        obs = E.expand_obs(boards, torch.float32)
        E.pack_obs(obs)
        E.pack_obs(E.expand_obs(boards, torch.bool))
        E.unpack_status(status)
        E.row_table_lookup(torch.arange(0, 65536, 7, dtype=torch.int32, device=dev).to(torch.int16))
    for n in (1, 5, 255, 6143, 6144, 6145, 20001):
        r = torch.rand(n, device=dev)
        v = torch.rand(n, device=dev)
        d = (torch.rand(n, device=dev) < 0.01).to(torch.uint8)
        for entry in ("g2048_gae_flat", "g2048_gae_flat_pipelined", "g2048_gae_flat_tiled", "g2048_gae_flat_v1", "g2048_gae_flat_scan"):
            adv, ret, mom = E.gae_flat(r, v, d, 0.99, 0.95, entry=entry)
            E.gae_flat(r, v, d, 0.99, 0.95, want_moments=False, entry=entry)
        E.normalize_(adv, mom, 1)
        E.normalize_(ret, mom, 3)
        if n > 8:
            E.gae_flat(r[3:], v[3:], d[3:], 0.99, 0.95)  # views that are not 16-byte aligned
    t_steps, n_envs = 37, 301
    rr = torch.rand(t_steps, n_envs, device=dev)
    vv = torch.rand(t_steps, n_envs, device=dev)
    mm = (torch.rand(t_steps, n_envs, device=dev) < 0.05).to(torch.uint8) << 6
    E.gae_time_major(rr, vv, mm, t_steps, n_envs, None, 0.99, 0.95)
    E.gae_time_major(rr, vv, mm, t_steps, n_envs, torch.rand(n_envs, device=dev), 0.99, 0.95)
    E.row_moments(torch.rand(3, 1001, dtype=torch.float64, device=dev))
    packed = dict(boards=torch.randint(0, 1 << 62, (5000,), dtype=torch.int64, device=dev),
                  meta=torch.randint(0, 127, (5000,), dtype=torch.uint8, device=dev),
                  log_probs=torch.rand(5000, device=dev), values=torch.rand(5000, device=dev))
    idx = torch.randperm(5000, device=dev)[:777].contiguous()
    for dt in (torch.float32, torch.bfloat16):
        E.gather_minibatch(idx, packed, torch.rand(5000, device=dev), torch.rand(5000, device=dev), obs_dtype=dt)
    for shape, dt in (((333, 41, 3), torch.uint8), ((333, 41, 5), torch.float32), ((333, 41, 4, 4), torch.float32)):
        term = torch.rand(333, 41, device=dev) < 0.05
        lens = E.first_done_rows(term)
        offs = E.exclusive_scan(lens)
        total = int(offs[-1])
        dst = T.empty((total, *shape[2:]), dtype=dt, device=dev)
        src = torch.zeros(shape, dtype=dt, device=dev)
        row_bytes = src[0, 0].numel() * src.element_size()
        if total:
            E.N.call("g2048_compact_rows", E.N.ptr(src), 333, 41, row_bytes, E.N.ptr(lens), E.N.ptr(offs), 0, E.N.ptr(dst), E.N.stream_ptr())
    E.random_subset(5000, 777, (1, 2), dev, first=13)
    E.random_subset(1, 1, 9, dev)
    eb = packed["boards"][:301].contiguous()
    for dt, d_model in ((torch.float32, 256), (torch.bfloat16, 256), (torch.float32, 12), (torch.bfloat16, 8)):
        table = torch.randn(31, d_model, device=dev).to(dt)
        for entry in ("g2048_embed_boards", "g2048_embed_boards_bulk", "g2048_embed_boards_plain"):
            E.embed_boards(eb, table, entry=entry)
            E.embed_boards(packed["boards"], table, idx, entry=entry)
        E.embed_boards_grad(eb, torch.randn(301, 16, d_model, device=dev).to(dt))
        E.embed_boards_grad(packed["boards"], torch.randn(777, 16, d_model, device=dev).to(dt), idx)
    sink = T.empty(4 * 64, dtype=torch.int32, device=dev)
    E.int_peak_probe(4, 64, 10, sink)
    E.play_host(0, 3, 500, 1)
    E.play_host(1, 3, 500, 0, pinned=True)
    E.play_packed(0, E.chain_advance(E.words_tensor([0, 5], dev), 1, 1 + 2 * 700), 777, 100, 300, 1)
    E.gae_host(np.random.rand(7000).astype(np.float32), np.random.rand(7000).astype(np.float32),
               (np.random.rand(7000) < 0.01).astype(np.uint8), 0.99, 0.95, True)
    runner = g2048.BatchRunner(1, g2048.act_randomly)
    runner.run_flat_batch(333)
    ro = runner.run_packed_batch(50)
    buf = g2048.RolloutBuffer(31, 16, 4)
    buf.store_packed(ro)
    buf.get_buffer_data()
    for _ in g2048.DevicePPOBatches(buf.get_packed(), batch_size=64):
        pass
    torch.cuda.synchronize()
