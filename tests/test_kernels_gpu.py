"""Parity of the CUDA kernels (through the C ABI) against the oracle.  Needs a B200: -m gpu."""
import json

import numpy as np
import pytest
import torch

from oracle import c_oracle as CO
from oracle import pgx2048_oracle as O

pytestmark = pytest.mark.gpu

MODES = [0, 1]  # original, partitionable


@pytest.fixture(scope="module")
def E():
    from g2048 import engine

    assert torch.cuda.is_available()
    return engine


def dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def u32(a):
    return dev(np.asarray(a, np.uint32).view(np.int32))


def mask_bits(status):
    return ((status[:, None] >> np.arange(4)) & 1).astype(bool)


def random_boards(rng, n):
    boards = rng.integers(0, 8, (n, 16)) * (rng.random((n, 16)) < 0.7)
    k = n // 20
    boards[:k] = rng.integers(1, 6, (k, 16))  # full boards: terminal / illegal-on-full paths
    boards[k : 2 * k] = rng.integers(0, 3, (k, 16))
    boards[2 * k : 2 * k + 50] = rng.integers(9, 14, (50, 16))
    return boards


def state_of(boards):
    masks = O.exact_legal(boards)
    done = ~masks.any(axis=1)
    masks = np.where(done[:, None], True, masks)
    status = ((masks * (1 << np.arange(4))).sum(axis=1) | np.where(done, 16, 0)).astype(np.uint8)
    return masks, done, status


# ----------------------------------------------------------------------------------------- RNG
def test_threefry_kats(E):
    keys = u32([[0, 0], [0xFFFFFFFF, 0xFFFFFFFF], [0x13198A2E, 0x03707344]])
    ctrs = u32([[0, 0], [0xFFFFFFFF, 0xFFFFFFFF], [0x243F6A88, 0x85A308D3]])
    out = E.words_numpy(E.threefry2x32(keys, ctrs))
    want = [[0x6B200159, 0x99BA4EFE], [0x1CB996FC, 0xBB002BE7], [0xC4923A9C, 0x483DF7A0]]
    np.testing.assert_array_equal(out, np.array(want, np.uint32))


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("n", [1, 2, 3, 4, 7, 1000, 4097])
def test_split_keys_match_oracle(E, mode, n):
    sub = [0xDEADBEEF, 0x12345678]
    got = E.words_numpy(E.split_keys(u32(sub), n, 0, n, mode))
    np.testing.assert_array_equal(got, CO.split(sub, n, mode))
    if n > 4:  # a shard in the middle gives the same keys (global indices)
        lo, m = n // 3, n // 2
        part = E.words_numpy(E.split_keys(u32(sub), n, lo, m, mode))
        np.testing.assert_array_equal(part, got[lo : lo + m])


@pytest.mark.parametrize("mode", MODES)
def test_chain_matches_oracle(E, mode):
    key = u32([0, 42])
    subs = E.words_numpy(E.chain_advance(key, mode, 301))
    want_key, want = CO.chain([0, 42], mode, 301)
    np.testing.assert_array_equal(subs, want)
    np.testing.assert_array_equal(E.words_numpy(key), want_key)


# ----------------------------------------------------------------------------------------- env
@pytest.mark.parametrize("mode", MODES)
def test_env_init_matches_oracle(E, mode):
    n = 20000
    sub = [7, 11]
    boards, status = E.env_init(u32(sub), n, 0, n, mode)
    wb, wm = CO.env_init(CO.split(sub, n, mode), mode)
    np.testing.assert_array_equal(E.boards_numpy(boards), wb)
    st = status.cpu().numpy()
    np.testing.assert_array_equal(mask_bits(st), wm.astype(bool))
    assert not (st & 0x30).any()
    # explicit-key form (jax.vmap(env.init)(keys))
    keys = CO.split(sub, n, mode)
    b2, s2 = E.env_init(u32(keys), 0, 0, n, mode)
    assert torch.equal(b2, boards) and torch.equal(s2, status)


def test_known_init_state_original_mode(E):
    # SURVEY Appendix A.5, pinned by the first SVG frames
    key = u32([0, 0])
    subs = E.chain_advance(key, 0, 1)
    boards, status = E.env_init(subs[0], 4, 0, 4, 0)
    b = E.boards_numpy(boards)
    want = [{8: 1, 12: 1}, {1: 1, 14: 1}, {8: 1, 10: 2}, {8: 1, 12: 1}]
    for e in range(4):
        assert {int(i): int(b[e, i]) for i in np.nonzero(b[e])[0]} == want[e]
    np.testing.assert_array_equal(mask_bits(status.cpu().numpy()), [[0, 1, 1, 1], [1, 1, 1, 1], [1, 1, 1, 1], [0, 1, 1, 1]])


@pytest.mark.parametrize("mode", MODES)
def test_env_step_bit_exact_on_a_million_random_triples(E, mode):
    rng = np.random.default_rng(100 + mode)
    n = 1_000_000
    boards = random_boards(rng, n)
    masks, done, status = state_of(boards)
    actions = rng.integers(0, 4, n).astype(np.int32)  # includes illegal actions
    sub = [0xABCDEF01, 0x2048]
    keys = CO.split(sub, n, mode)
    wb, wm, wd, wr = CO.env_step(boards, masks, done, actions, keys, mode)
    d_b, d_s = dev(E.pack_boards(boards)), dev(status)
    rew = E.env_step(d_b, d_s, dev(actions), u32(sub), n, 0, mode)
    np.testing.assert_array_equal(E.boards_numpy(d_b), wb)
    st = d_s.cpu().numpy()
    np.testing.assert_array_equal(mask_bits(st), wm.astype(bool))
    np.testing.assert_array_equal((st >> 4) & 1, wd)
    np.testing.assert_array_equal(rew.cpu().numpy(), wr)
    assert (wr == -1).sum() > 100 and wd.sum() > 100 and (wr > 0).sum() > 1000


def test_env_step_with_given_draws_is_bit_exact(E):
    """Board, reward, mask and done given identical actions and spawn draws (north_star wording)."""
    rng = np.random.default_rng(3)
    n = 300_000
    boards = random_boards(rng, n)
    masks, done, status = state_of(boards)
    actions = rng.integers(0, 4, n).astype(np.int32)
    bits_pos = rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32)
    bits_val = rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32)
    bits_pos[:64] = [0, 0xFFFFFFFF, 0x1FF, 0x200] * 16  # extremes of (1 - u)
    bits_val[:8] = [0, 0xFFFFFFFF, 0x19999800, 0x19999A00, 0x199999FF, 0x19999C00, 0x19999600, 0x1999A000]  # around 0.9f
    u_pos = O._bits_to_unit_float(bits_pos)
    u_val = O._bits_to_unit_float(bits_val)
    wb, wm, wd, wr = CO.env_step_given(boards, masks, done, actions, u_pos, u_val)
    d_b, d_s = dev(E.pack_boards(boards)), dev(status)
    rew = E.env_step_draws(d_b, d_s, dev(actions), u32(bits_pos), u32(bits_val))
    np.testing.assert_array_equal(E.boards_numpy(d_b), wb)
    st = d_s.cpu().numpy()
    np.testing.assert_array_equal(mask_bits(st), wm.astype(bool))
    np.testing.assert_array_equal((st >> 4) & 1, wd)
    np.testing.assert_array_equal(rew.cpu().numpy(), wr)
    # and against the numpy restatement on a slice (two independent oracles)
    sl = slice(0, 5000)
    st0 = O.State(boards[sl].astype(np.int32), masks[sl], np.zeros((5000, 1), np.float32), done[sl], np.zeros(5000, bool))
    ref = O.env_step_given(st0, actions[sl], u_pos[sl], u_val[sl])
    np.testing.assert_array_equal(E.boards_numpy(d_b)[sl], ref.board)
    np.testing.assert_array_equal(rew.cpu().numpy()[sl], ref.rewards[:, 0])


def test_step_hand_built_edge_boards(E):
    rows = [
        ([1, 1, 1, 1] + [0] * 12, 0, [2, 2, 0, 0] + [0] * 12, 8),
        ([1, 1, 2, 0] + [0] * 12, 2, [0, 0, 2, 2] + [0] * 12, 4),
        ([3, 0, 0, 0, 3, 0, 0, 0, 0, 0, 0, 0, 3, 0, 0, 0], 1, [4, 0, 0, 0, 3] + [0] * 11, 16),
        ([3, 0, 0, 0, 3, 0, 0, 0, 0, 0, 0, 0, 3, 0, 0, 0], 3, [0] * 8 + [3, 0, 0, 0, 4, 0, 0, 0], 16),
        ([14, 14, 0, 0] + [0] * 12, 0, [15, 0, 0, 0] + [0] * 12, 1 << 15),
    ]
    boards = np.array([r[0] for r in rows])
    actions = np.array([r[1] for r in rows], np.int32)
    masks, done, status = state_of(boards)
    d_b, d_s = dev(E.pack_boards(boards)), dev(status)
    # draws that put a 2 in the LAST empty cell: u = 0 -> (1-u) = 1 -> r = n_empty
    rew = E.env_step_draws(d_b, d_s, dev(actions), u32(np.zeros(5, np.uint32)), u32(np.full(5, 0xFFFFFFFF, np.uint32)))
    got = E.boards_numpy(d_b)
    for i, (_, _, want, r) in enumerate(rows):
        moved = got[i].copy()
        last_empty = np.nonzero(np.array(want) == 0)[0][-1]
        assert moved[last_empty] == 1
        moved[last_empty] = 0
        assert moved.tolist() == want, i
        assert rew[i].item() == r
    # 2^15 + 2^15 cannot be stored in a nibble: the overflow flag must be raised
    b = np.zeros((1, 16), np.int64)
    b[0, :2] = 15
    masks, done, status = state_of(b)
    d_b, d_s = dev(E.pack_boards(b)), dev(status)
    E.env_step_draws(d_b, d_s, dev(np.zeros(1, np.int32)), u32([0]), u32([0]))
    assert d_s.item() & 0x20


@pytest.mark.parametrize("mode", MODES)
def test_frozen_env_and_terminal_mask(E, mode):
    full = np.array([[1, 2, 1, 2, 2, 1, 2, 1, 1, 2, 1, 2, 2, 1, 2, 1]])  # no legal move
    masks, done, status = state_of(full)
    assert done[0] and status[0] == 0x1F
    d_b, d_s = dev(E.pack_boards(full)), dev(status)
    rew = E.env_step(d_b, d_s, dev(np.array([2], np.int32)), u32([1, 2]), 1, 0, mode)
    assert rew.item() == 0.0 and d_s.item() == 0x1F
    np.testing.assert_array_equal(E.boards_numpy(d_b), full)


# ----------------------------------------------------------------------------------------- policies
@pytest.mark.parametrize("mode", MODES)
def test_act_matches_oracle(E, mode):
    rng = np.random.default_rng(17)
    n = 200_000
    masks = rng.random((n, 4)) < 0.6
    masks[masks.sum(1) == 0, 1] = True
    status = (masks * (1 << np.arange(4))).sum(axis=1).astype(np.uint8)
    sub = [5, 6]
    keys = CO.split(sub, n, mode)
    wa, wlp = CO.act(keys, masks, CO.RANDOM, mode)
    a, lp = E.act(E.POLICY_RANDOM, dev(status), u32(sub), n, 0, mode)
    np.testing.assert_array_equal(a.cpu().numpy(), wa)
    np.testing.assert_allclose(lp.cpu().numpy(), wlp, rtol=1e-6, atol=0)
    a, lp = E.act(E.POLICY_DRUL, dev(status), u32(sub), n, 0, mode)
    assert lp is None
    np.testing.assert_array_equal(a.cpu().numpy(), CO.act(None, masks, CO.DRUL, mode)[0])
    # numpy oracle with the full gumbel formula on a slice
    a_np, lp_np = O.act_randomly((keys[:5000, 0], keys[:5000, 1]), masks[:5000], mode)
    np.testing.assert_array_equal(wa[:5000], a_np)


def test_act_kats_of_the_reference_tests(E):
    # tests/actions/test_act_drul.py:40-48, tests/actions/test_act_randomly.py:42-53
    status = dev(np.array([0b0101], np.uint8))
    for seed in range(16):
        for mode in MODES:
            a, lp = E.act(E.POLICY_RANDOM, status, u32([[0, seed]]), 0, 0, mode)
            assert a.item() in (0, 2)
            assert abs(lp.item() - np.log(0.5)) < 1e-6
    a, _ = E.act(E.POLICY_DRUL, status, u32([[0, 0]]), 0, 0, 1)
    assert a.item() == 2


@pytest.mark.parametrize("mode", MODES)
def test_sample_logits_matches_oracle(E, mode, golden_ppo):
    rng = np.random.default_rng(23)
    n = 100_000
    logits = (rng.standard_normal((n, 4)) * 3).astype(np.float32)
    masks = rng.random((n, 4)) < 0.6
    masks[masks.sum(1) == 0, 3] = True
    status = (masks * (1 << np.arange(4))).sum(axis=1).astype(np.uint8)
    sub = [99, 1234]
    keys = CO.split(sub, n, mode)
    masked = O.mask_logits(logits, masks)
    wa, wlp = O.act_from_logits((keys[:, 0], keys[:, 1]), masked, mode, sample=True)
    a, lp, ent = E.sample_logits(dev(logits), dev(status), True, True, u32(sub), n, 0, mode, want_entropy=True)
    a, lp = a.cpu().numpy(), lp.cpu().numpy()
    # gumbel-max depends on the last ulp of logf: near-ties may resolve differently
    agree = a == wa
    assert agree.mean() > 0.9999
    np.testing.assert_allclose(lp[agree], wlp[agree], rtol=1e-5, atol=1e-6)
    assert masks[np.arange(n), a].all()
    # log-prob / entropy of given actions vs torch.distributions (ppo_agent.py:182-189)
    t_logits = torch.from_numpy(masked)
    dist = torch.distributions.Categorical(logits=t_logits)
    np.testing.assert_allclose(lp, dist.log_prob(torch.from_numpy(a).long()).numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(ent.cpu().numpy(), dist.entropy().numpy(), rtol=1e-5, atol=1e-6)
    # argmax path
    a2, lp2, _ = E.sample_logits(dev(logits), dev(status), True, False, None, n, 0, mode)
    np.testing.assert_array_equal(a2.cpu().numpy(), masked.argmax(1))


def test_evaluate_logits_matches_reference_fixture(E, golden_ppo):
    g = golden_ppo
    bits = (g["lp_masks"] * (1 << np.arange(4))).sum(axis=1).astype(np.uint8)
    for pre in ("lp", "lp2"):
        lp, ent = E.evaluate_logits(dev(g[f"{pre}_raw_logits"]), dev(bits), True, dev(g[f"{pre}_actions"]))
        np.testing.assert_allclose(lp.cpu().numpy(), g[f"{pre}_log_probs"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(ent.cpu().numpy(), g[f"{pre}_entropy"], rtol=1e-5, atol=1e-6)


# ----------------------------------------------------------------------------------------- fused loops
@pytest.mark.parametrize("entry", ["g2048_play", "g2048_play_tables"])
@pytest.mark.parametrize("policy,name", [(1, "drul"), (0, "random")])
def test_play_reproduces_golden_svg(E, golden_svg, golden_hist, policy, name, entry):
    want = golden_hist["svg_seed0_batch4_original_mode"][name]
    key = u32([0, 0])
    subs = E.chain_advance(key, 0, 1 + 2 * 1024)
    out = E.play(policy, subs, 4, 0, 4, 0, entry=entry)
    assert out["lengths"].cpu().tolist() == want["lengths"]
    assert out["scores"].cpu().tolist() == want["scores"]
    np.testing.assert_array_equal(E.boards_numpy(out["final_boards"]), golden_svg[f"{name}_boards"][-1])
    st = E.play_stats_dict(out["stats"])
    assert st["episodes"] == 4 and st["longest"] == want["loop_steps"] and st["cut_short"] == 0
    assert st["env_steps"] == sum(want["lengths"]) and st["score_sum"] == sum(want["scores"])


@pytest.mark.parametrize("policy,name", [(0, "random"), (1, "drul")])
def test_play_reproduces_png_histograms(E, golden_hist, policy, name):
    hist = {}
    for seed in golden_hist["seeds"]:
        key = u32(list(E.key_words(seed)))
        subs = E.chain_advance(key, 1, 1 + 2 * 2048)
        st = E.play_stats_dict(E.play(policy, subs, 100, 0, 100, 1)["stats"])
        for k, v in st["max_tile_hist"].items():
            hist[k] = hist.get(k, 0) + v
    got = {str(k): round(v / 10, 1) for k, v in sorted(hist.items())}
    assert got == golden_hist[name]


PLAY_ENTRIES = ["g2048_play", "g2048_play_tables", "g2048_play_swar", "g2048_play_v1"]


@pytest.mark.parametrize("entry", PLAY_ENTRIES)
@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("policy", [0, 1])
def test_play_matches_oracle_and_is_shard_invariant(E, mode, policy, entry):
    n = 40000 if entry == "g2048_play" else 20000  # the dispatcher switches to the table kernel at 32768 envs
    seed = 1234
    want = CO.play(seed, n, policy, mode, max_steps=2048)
    key = u32(list(E.key_words(seed)))
    subs = E.chain_advance(key, mode, 1 + 2 * 2048)
    out = E.play(policy, subs, n, 0, n, mode, entry=entry)
    np.testing.assert_array_equal(E.boards_numpy(out["final_boards"]), want["final_boards"])
    np.testing.assert_array_equal(out["lengths"].cpu().numpy(), want["lengths"])
    np.testing.assert_array_equal(out["scores"].cpu().numpy(), want["scores"])
    st = E.play_stats_dict(out["stats"])
    assert st["env_steps"] == int(want["lengths"].sum()) and st["longest"] == want["longest"]
    tiles = 1 << want["final_boards"].max(axis=1).astype(np.int64)
    assert st["tile_sum"] == int(tiles.sum()) and st["tile_sq_sum"] == int((tiles**2).sum())
    # shard [lo, lo+m) of the same global batch == the slice (what each GPU computes)
    lo, m = 7001, 5000
    part = E.play(policy, subs, n, lo, m, mode, entry=entry)
    np.testing.assert_array_equal(E.boards_numpy(part["final_boards"]), want["final_boards"][lo : lo + m])
    np.testing.assert_array_equal(part["lengths"].cpu().numpy(), want["lengths"][lo : lo + m])


def test_play_host_entry_point(E, golden_hist):
    want = golden_hist["svg_seed0_batch4_original_mode"]["drul"]
    out = E.play_host(E.POLICY_DRUL, 0, 4, 0, key=np.array([0, 0], np.uint32))
    assert out["lengths"].tolist() == want["lengths"] and out["scores"].tolist() == want["scores"]
    # the chain key the reference's runner would hold after the run: 1 + 2*T splits
    want_key, _ = CO.chain([0, 0], 0, 1 + 2 * want["loop_steps"])
    np.testing.assert_array_equal(out["key"], want_key)
    # a larger batch through the same host entry
    out = E.play_host(E.POLICY_DRUL, 5, 3000, 1)
    ref = CO.play(5, 3000, CO.DRUL, 1, max_steps=4096)
    np.testing.assert_array_equal(out["lengths"], ref["lengths"])


def test_play_host_pinned_results_are_written_by_the_kernel(E):
    """Pinned (mapped) result arrays take the zero-copy path of g2048_play_host: same results as pageable arrays and as
    the oracle, including a batch large enough for the table kernel and a second call into the same arrays."""
    for policy, seed, n, mode in ((E.POLICY_RANDOM, 11, 40000, 1), (E.POLICY_DRUL, 12, 3000, 0)):
        a = E.play_host(policy, seed, n, mode, pinned=True)
        b = E.play_host(policy, seed, n, mode, pinned=False)
        ref = CO.play(seed, n, policy, mode, max_steps=4096)
        for k in ("final_boards", "lengths", "scores"):
            np.testing.assert_array_equal(a[k], b[k], err_msg=k)
        np.testing.assert_array_equal(E.boards_numpy(torch.from_numpy(a["final_boards"].view(np.int64))), ref["final_boards"])
        np.testing.assert_array_equal(a["lengths"], ref["lengths"])
        np.testing.assert_array_equal(a["scores"], ref["scores"])
        np.testing.assert_array_equal(a["stats"], b["stats"])
        assert int(a["lengths"].sum()) == int(a["stats"][1])
        # the packed device entry and the packed host entry with a pageable record array (staged copy)
        subs = E.chain_advance(E.words_tensor(list(E.key_words(seed)), "cuda"), mode, 1 + 2 * 4096)
        rec = E.play_packed(policy, subs, n, 0, n, mode)["results"].cpu().numpy().view(E.EPISODE_RESULT).reshape(n)
        np.testing.assert_array_equal(rec, a["records"])
        pageable = np.zeros(n, E.EPISODE_RESULT)
        E.N.call("g2048_play_host_packed", policy, seed, None, n, 0, n, mode, pageable.ctypes.data, None)
        np.testing.assert_array_equal(pageable, a["records"])


def test_host_entries_with_multi_megabyte_pageable_arrays(E):
    """The staged copier splits the host-side memcpy of a chunk over several threads once it has a megabyte or more
    to move (g2048_hostcopy.cuh): results equal the zero-copy path's, element for element, also for sizes that are not
    a multiple of the piece or of the 16 MiB staging buffer."""
    n = 2_500_003  # boards: 20 MB (two staging chunks, the second one short), lengths / scores: 10 MB
    a = E.play_host(E.POLICY_RANDOM, 21, n, 1, pinned=True)
    b = E.play_host(E.POLICY_RANDOM, 21, n, 1, pinned=False)
    for k in ("final_boards", "lengths", "scores"):
        np.testing.assert_array_equal(a[k], b[k], err_msg=k)
    np.testing.assert_array_equal(a["stats"], b["stats"])
    # host-array GAE: 3 x 14 MB in through the same copier, 2 x 14 MB out
    rng = np.random.default_rng(3)
    m = 3_500_017
    r = (rng.integers(0, 64, m) * 4).astype(np.float32)
    v = rng.standard_normal(m).astype(np.float32)
    d = (rng.random(m) < 1 / 120).astype(np.uint8)
    adv, ret = E.gae_host(r, v, d, 0.99, 0.95, False)
    want_adv, want_ret, _ = E.gae_flat(torch.from_numpy(r).cuda(), torch.from_numpy(v).cuda(), torch.from_numpy(d).cuda(), 0.99, 0.95)
    np.testing.assert_array_equal(adv, want_adv.cpu().numpy())
    np.testing.assert_array_equal(ret, want_ret.cpu().numpy())


@pytest.mark.parametrize("entry", PLAY_ENTRIES)
def test_play_reports_cut_short(E, entry):
    key = u32([0, 1])
    subs = E.chain_advance(key, 1, 1 + 2 * 20)  # only 20 loop steps of keys
    st = E.play_stats_dict(E.play(0, subs, 64, 0, 64, 1, entry=entry)["stats"])
    assert st["cut_short"] == 64 and st["longest"] == 20


def test_row_tables_match_the_oracle_for_every_row(E):
    """All 65 536 entries of the shared-memory tables behind g2048_play_tables against the oracle's move."""
    rows = np.arange(65536, dtype=np.uint32)
    cells = np.stack([(rows >> (4 * c)) & 15 for c in range(4)], axis=1).astype(np.int64)
    left, flags = E.row_table_lookup(dev(rows.astype(np.uint16).view(np.int16)))
    left = left.cpu().numpy().view(np.uint16).astype(np.uint32)
    flags = flags.cpu().numpy()
    boards = np.zeros((65536, 16), np.int64)
    boards[:, :4] = cells
    ok = cells.max(axis=1) < 15  # 2^15 + 2^15 does not fit a nibble: those envs are flagged, not represented
    want_l, _ = O.move(boards[ok], np.zeros(ok.sum(), np.int64))
    want_r, _ = O.move(boards[ok], np.full(ok.sum(), 2))
    got_l = np.stack([(left[ok] >> (4 * c)) & 15 for c in range(4)], axis=1)
    np.testing.assert_array_equal(got_l, want_l[:, :4])
    np.testing.assert_array_equal(flags[ok] & 1, (want_l[:, :4] != cells[ok]).any(axis=1))
    np.testing.assert_array_equal((flags[ok] >> 2) & 1, (want_r[:, :4] != cells[ok]).any(axis=1))
    assert not (flags & ~np.uint8(5)).any()


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("policy,name", [(0, "random"), (1, "drul")])
def test_rollout_steps_records_match_numpy_oracle(E, mode, policy, name):
    n, seed, chunk = 64, 77, 50
    ref = O.rollout(O.KeyChain(seed, mode), n, name)
    t_ref = len(ref["states"])
    key = u32(list(E.key_words(seed)))
    subs = E.chain_advance(key, mode, 1 + 2 * 1024)
    boards, status = E.env_init(subs[0], n, 0, n, mode)
    counters = torch.zeros(4, dtype=torch.int64, device="cuda")
    recs = []
    t0 = 0
    while True:
        rb = torch.empty((chunk, n), dtype=torch.int64, device="cuda")
        rm = torch.empty((chunk, n), dtype=torch.uint8, device="cuda")
        rr = torch.empty((chunk, n), dtype=torch.float32, device="cuda")
        rl = torch.empty((chunk, n), dtype=torch.float32, device="cuda") if policy == 0 else None
        E.rollout_steps(policy, boards, status, subs[1 + 2 * t0 :], chunk, t0, n, 0, mode, rb, rm, rr, rl, counters)
        recs.append((rb, rm, rr, rl))
        t0 += chunk
        if counters[0].item() == n:
            break
    t_total = int(counters[1].item())
    assert t_total == t_ref
    rb = torch.cat([r[0] for r in recs])[:t_total]
    rm = torch.cat([r[1] for r in recs])[:t_total].cpu().numpy()
    rr = torch.cat([r[2] for r in recs])[:t_total].cpu().numpy()
    np.testing.assert_array_equal(E.boards_numpy(rb), np.stack(ref["boards"]))
    np.testing.assert_array_equal(rm & 3, np.stack(ref["actions"]))
    np.testing.assert_array_equal(((rm[..., None] >> (2 + np.arange(4))) & 1).astype(bool), np.stack(ref["masks"]))
    np.testing.assert_array_equal((rm >> 6) & 1, np.stack([s.terminated for s in ref["states"]]))
    np.testing.assert_array_equal(rr, np.stack([s.rewards[:, 0] for s in ref["states"]]))
    if policy == 0:
        rl = torch.cat([r[3] for r in recs])[:t_total].cpu().numpy()
        np.testing.assert_allclose(rl, np.stack(ref["log_probs"]), rtol=1e-6)
    length, score, _ = O.episode_summary(ref)
    assert int(counters[2].item()) == int(length.sum()) and int(counters[3].item()) == int(score.sum())
    np.testing.assert_array_equal(E.boards_numpy(boards), ref["states"][-1].board)


@pytest.mark.parametrize("mode", MODES)
def test_policy_step_equals_sample_then_step(E, mode):
    rng = np.random.default_rng(31)
    n = 50_000
    boards = random_boards(rng, n)
    masks, done, status = state_of(boards)
    logits = (rng.standard_normal((n, 4)) * 2).astype(np.float32)
    values = rng.standard_normal(n).astype(np.float32)
    sa, ss = u32([1, 2]), u32([3, 4])
    a_ref, lp_ref, _ = E.sample_logits(dev(logits), dev(status), True, True, sa, n, 0, mode)
    b_ref, s_ref = dev(E.pack_boards(boards)), dev(status)
    r_ref = E.env_step(b_ref, s_ref, a_ref, ss, n, 0, mode)
    b, s = dev(E.pack_boards(boards)), dev(status)
    rb = torch.empty(n, dtype=torch.int64, device="cuda")
    rm = torch.empty(n, dtype=torch.uint8, device="cuda")
    rr, rl, rv = (torch.empty(n, dtype=torch.float32, device="cuda") for _ in range(3))
    acts = torch.empty(n, dtype=torch.int32, device="cuda")
    E.policy_step(b, s, dev(logits), dev(values), True, True, False, sa, ss, n, 0, mode, rb, rm, rr, rl, rv, acts)
    assert torch.equal(b, b_ref) and torch.equal(s, s_ref) and torch.equal(rr, r_ref)
    assert torch.equal(acts, a_ref) and torch.equal(rl, lp_ref)
    assert torch.equal(rb, dev(E.pack_boards(boards))) and torch.equal(rv, dev(values))
    m = rm.cpu().numpy()
    np.testing.assert_array_equal(m & 3, a_ref.cpu().numpy())
    np.testing.assert_array_equal((m >> 2) & 15, status & 15)
    np.testing.assert_array_equal((m >> 6) & 1, (s_ref.cpu().numpy() >> 4) & 1)
    # sampled actions are legal wherever the env was alive (masked logits)
    alive = ~done
    assert masks[np.arange(n), a_ref.cpu().numpy()][alive].all()


def test_policy_step_auto_reset_follows_pgx_wrapper(E):
    mode, n = 1, 65536
    rng = np.random.default_rng(37)
    boards = rng.integers(1, 4, (n, 16))  # nearly dead boards: many terminate on this step
    masks, done, status = state_of(boards)
    live = ~done
    logits = np.zeros((n, 4), np.float32)
    sa, ss = [10, 20], [30, 40]
    b, s = dev(E.pack_boards(boards)), dev(status)
    rr = torch.empty(n, dtype=torch.float32, device="cuda")
    rm = torch.empty(n, dtype=torch.uint8, device="cuda")
    acts = torch.empty(n, dtype=torch.int32, device="cuda")
    E.policy_step(b, s, dev(logits), None, True, True, True, u32(sa), u32(ss), n, 0, mode, None, rm, rr, None, None, acts)
    # oracle: k1, k2 = split(step_key); step with k1; if done -> init(k2) keeping done/reward
    step_keys = CO.split(ss, n, mode)
    k0, k1 = O.split((step_keys[:, 0], step_keys[:, 1]), 2, mode)
    key1 = np.stack([k0[:, 0], k1[:, 0]], axis=1)
    key2 = np.stack([k0[:, 1], k1[:, 1]], axis=1)
    a = acts.cpu().numpy()
    # a previously finished env is un-done first (its mask is then the exact mask of its board: none legal)
    wb, wm, wd, wr = CO.env_step(boards, O.exact_legal(boards), np.zeros(n, bool), a, key1, mode)
    ib, im = CO.env_init(key2, mode)
    wb = np.where(wd[:, None].astype(bool), ib, wb)
    wm = np.where(wd[:, None].astype(bool), im, wm)
    sel = live  # envs that were alive before the step took a legal action
    np.testing.assert_array_equal(E.boards_numpy(b)[sel], wb[sel])
    st = s.cpu().numpy()
    np.testing.assert_array_equal(((st >> 4) & 1)[sel], wd[sel])
    np.testing.assert_array_equal(mask_bits(st)[sel], wm[sel].astype(bool))
    np.testing.assert_array_equal(rr.cpu().numpy()[sel], wr[sel])
    assert wd[sel].sum() > 10


@pytest.mark.parametrize("mode", MODES)
def test_auto_reset_undoes_a_finished_env_before_stepping_it(E, mode):
    """pgx.experimental.auto_reset, the branch SURVEY A.6 lists first: a state that finished on the previous step was
    replaced by a fresh one but still carries terminated=True; the wrapper clears the flag (and the rewards) and then
    steps it like any other env.  Two consecutive steps: the envs that finish on the first are the ones entering the
    second with the flag set -- their second step must be the oracle's plain env.step on the fresh board."""
    n = 65536
    rng = np.random.default_rng(41)
    boards = rng.integers(1, 4, (n, 16))  # nearly dead boards: many terminate on the first step
    masks, done, status = state_of(boards)
    keep = ~done  # start from live envs only
    boards, status = boards[keep], status[keep]
    n = boards.shape[0]
    b, s = dev(E.pack_boards(boards)), dev(status)
    logits = dev(np.zeros((n, 4), np.float32))
    rr = torch.empty(n, dtype=torch.float32, device="cuda")
    rm = torch.empty(n, dtype=torch.uint8, device="cuda")
    acts = torch.empty(n, dtype=torch.int32, device="cuda")
    E.policy_step(b, s, logits, None, True, True, True, u32([10, 20]), u32([30, 40]), n, 0, mode, None, rm, rr, None, None, acts)
    st1 = s.cpu().numpy()
    finished = (st1 & 16) != 0
    assert finished.sum() > 10
    boards1 = E.boards_numpy(b)  # fresh boards (two tiles) where `finished`
    assert ((boards1[finished] != 0).sum(axis=1) == 2).all()
    np.testing.assert_array_equal(mask_bits(st1)[finished], O.exact_legal(boards1[finished]))  # the fresh state's own mask
    pre2 = b.clone()
    rb2 = torch.empty(n, dtype=torch.int64, device="cuda")
    E.policy_step(b, s, logits, None, True, True, True, u32([11, 21]), u32([31, 41]), n, 0, mode, rb2, rm, rr, None, None, acts)
    assert torch.equal(rb2, pre2)  # the record holds the fresh board, not the dead one
    a2 = acts.cpu().numpy()
    assert mask_bits(st1)[np.arange(n), a2].all()  # legal under the mask the env showed, finished or not
    step_keys = CO.split([31, 41], n, mode)
    k0, k1 = O.split((step_keys[:, 0], step_keys[:, 1]), 2, mode)
    key1 = np.stack([k0[:, 0], k1[:, 0]], axis=1)
    key2 = np.stack([k0[:, 1], k1[:, 1]], axis=1)
    wb, wm, wd, wr = CO.env_step(boards1, mask_bits(st1), np.zeros(n, bool), a2, key1, mode)  # every env un-done first
    ib, im = CO.env_init(key2, mode)
    wb = np.where(wd[:, None].astype(bool), ib, wb)
    wm = np.where(wd[:, None].astype(bool), im, wm)
    st2 = s.cpu().numpy()
    np.testing.assert_array_equal(E.boards_numpy(b), wb)
    np.testing.assert_array_equal((st2 >> 4) & 1, wd)
    np.testing.assert_array_equal(mask_bits(st2), wm.astype(bool))
    np.testing.assert_array_equal(rr.cpu().numpy(), wr)
    np.testing.assert_array_equal((rm.cpu().numpy() >> 6) & 1, wd)
    assert (wr[finished] >= 0).all() and not wd[finished].any()  # a two-tile board never ends on its first step


# ----------------------------------------------------------------------------------------- records
@pytest.mark.parametrize("entry", ["g2048_expand_obs", "g2048_expand_obs_v1"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bool, torch.bfloat16])
def test_expand_obs_matches_oracle(E, dtype, entry):
    rng = np.random.default_rng(41)
    for n in (1, 2, 3, 5, 31, 257):  # ragged image groups of the bulk-store kernel
        b = rng.integers(0, 16, (n, 16))
        got = E.expand_obs(dev(E.pack_boards(b)), dtype, entry=entry)
        np.testing.assert_array_equal(got.float().cpu().numpy(), O.observe(b).reshape(n, 16, 31).astype(np.float32))
    n = 70001
    boards = rng.integers(0, 16, (n, 16))
    boards[0] = 0
    boards[1] = 15
    out = E.expand_obs(dev(E.pack_boards(boards)), dtype, entry=entry)
    want = O.observe(boards).reshape(n, 16, 31)
    np.testing.assert_array_equal(out.float().cpu().numpy(), want.astype(np.float32))
    # time-major records -> env-major (B, T) stacking of batch_runner.py:138
    t_steps, b = 7, 13
    rec = rng.integers(0, 12, (t_steps, b, 16))
    out = E.expand_obs(dev(E.pack_boards(rec)), dtype, rows=t_steps, n_cols=b, entry=entry)
    want = O.observe(rec.transpose(1, 0, 2).reshape(-1, 16)).reshape(b * t_steps, 16, 31)
    np.testing.assert_array_equal(out.float().cpu().numpy(), want.astype(np.float32))


def test_unpack_and_compact_match_reference_rollout_buffer(E, golden_ppo):
    g = golden_ppo
    b, t = g["rb_terminations"].shape
    meta = (g["rb_actions"] | ((g["rb_masks"] * (1 << np.arange(4))).sum(-1) << 2) | (g["rb_terminations"] << 6)).astype(np.uint8)
    rec = dict(
        boards=dev(E.pack_boards(g["rb_boards"]).T.copy()), meta=dev(meta.T.copy()), rewards=dev(g["rb_rewards"].T.copy()),
        log_probs=dev(g["rb_log_probs"].T.copy()), values=dev(g["rb_values"].T.copy()),
    )
    un = E.unpack_records(rec["meta"], rec["rewards"], rec["log_probs"], rec["values"], t, b)
    np.testing.assert_array_equal(un["actions"].cpu().numpy(), g["rb_actions"])
    np.testing.assert_array_equal(un["action_masks"].cpu().numpy(), g["rb_masks"])
    np.testing.assert_array_equal(un["terminations"].cpu().numpy(), g["rb_terminations"])
    np.testing.assert_array_equal(un["rewards"].cpu().numpy(), g["rb_rewards"])
    np.testing.assert_array_equal(un["values"].cpu().numpy(), g["rb_values"])
    lengths = E.episode_lengths(rec["meta"], t, b)
    assert lengths.cpu().tolist() == [5, 11, 1, 4, 0, 11, 7]
    offsets = E.exclusive_scan(lengths)
    assert offsets.cpu().tolist() == [0, 5, 16, 17, 21, 21, 32, 39]
    total = int(offsets[-1].item())
    fb = torch.empty(total, dtype=torch.int64, device="cuda")
    fm = torch.empty(total, dtype=torch.uint8, device="cuda")
    fr, fl, fv = (torch.empty(total, dtype=torch.float32, device="cuda") for _ in range(3))
    E.compact_records(rec["boards"], rec["meta"], rec["rewards"], rec["log_probs"], rec["values"], t, b, lengths, offsets, 0, fb, fm, fr, fl, fv)
    # the reference stored this batch and then its first three envs again (make_golden_ppo.py)
    k = total
    np.testing.assert_array_equal(fr.cpu().numpy(), g["rb_out_rewards"][:k])
    np.testing.assert_array_equal(fv.cpu().numpy(), g["rb_out_values"][:k])
    np.testing.assert_array_equal(fl.cpu().numpy(), g["rb_out_log_probs"][:k])
    obs = E.expand_obs(fb, torch.float32)
    np.testing.assert_array_equal(obs.cpu().numpy(), g["rb_out_observations"][:k])
    onehot, masks, term = E.unpack_flat_meta(fm)
    np.testing.assert_array_equal(onehot.cpu().numpy(), g["rb_out_actions"][:k])
    np.testing.assert_array_equal(masks.cpu().numpy(), g["rb_out_action_masks"][:k])
    np.testing.assert_array_equal(term.cpu().numpy(), g["rb_out_terminations"][:k])


@pytest.mark.parametrize("t_steps", [1, 15, 16, 17, 33, 200])
def test_episode_lengths_first_done(E, t_steps):
    """first done + 1 per env of time-major meta bytes (bit 6), 0 for an env that never terminates; the kernel reads
    16 steps per round trip, so the step counts straddle its batch size."""
    rng = np.random.default_rng(t_steps)
    n = 1000
    meta = rng.integers(0, 64, (t_steps, n)).astype(np.uint8)  # bits 0-5 arbitrary, bit 6 clear
    first = rng.integers(0, t_steps + 1, n)  # t_steps = never
    for e in range(n):
        if first[e] < t_steps:
            meta[first[e]:, e] |= (rng.random(t_steps - first[e]) < 0.5).astype(np.uint8) << 6  # later flags may be anything
            meta[first[e], e] |= 0x40
    want = np.where(first < t_steps, first + 1, 0).astype(np.uint32)
    got = E.episode_lengths(dev(meta), t_steps, n).cpu().numpy().view(np.uint32)
    np.testing.assert_array_equal(got, want)


@pytest.mark.parametrize("n", [0, 1, 31, 2047, 2048, 2049, 4096, 100_003, 2048 * 1024, 2048 * 1025 + 7, (1 << 24) + 5])
def test_exclusive_scan_sizes(E, n):
    """Chunked scan (2 048 elements per CTA, chunk sums parked in the output): empty, single element, chunk boundaries,
    exactly 1 024 chunks (one pass of the base kernel) and beyond, C5's 2^24; values up to 2^32 - 1 sum in 64 bits."""
    rng = np.random.default_rng(43 + n % 97)
    x = rng.integers(0, 400, n).astype(np.int64)
    if n > 3:
        x[[0, n // 2, n - 1]] = (1 << 32) - 1
    out = E.exclusive_scan(dev(x.astype(np.uint32).view(np.int32))).cpu().numpy()
    assert out.shape == (n + 1,)
    np.testing.assert_array_equal(out, np.concatenate([[0], np.cumsum(x)]))


# ----------------------------------------------------------------------------------------- GAE
TAGS = ["default", "short_eps", "undiscounted", "lowlam", "open_tail", "two"]


GAE_ENTRIES = ["g2048_gae_flat", "g2048_gae_flat_pipelined", "g2048_gae_flat_tiled", "g2048_gae_flat_v1"]


@pytest.mark.parametrize("entry", GAE_ENTRIES)
@pytest.mark.parametrize("tag", TAGS)
def test_gae_flat_is_bit_exact_vs_reference_fixture(E, golden_ppo, tag, entry):
    g = golden_ppo
    gamma, lam = g[f"gae_{tag}_params"]
    adv, ret, mom = E.gae_flat(dev(g[f"gae_{tag}_rewards"]), dev(g[f"gae_{tag}_values"]), dev(g[f"gae_{tag}_dones"].astype(np.uint8)), gamma, lam, entry=entry)
    np.testing.assert_array_equal(adv.cpu().numpy(), g[f"gae_{tag}_adv"])
    np.testing.assert_array_equal(ret.cpu().numpy(), g[f"gae_{tag}_ret"])
    # tolerance stated by north_star: 1e-5 relative in fp32
    E.normalize_(adv, mom, 1)
    E.normalize_(ret, mom, 3)
    np.testing.assert_allclose(adv.cpu().numpy(), g[f"gae_{tag}_adv_norm"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(ret.cpu().numpy(), g[f"gae_{tag}_ret_norm"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("entry", GAE_ENTRIES)
@pytest.mark.parametrize("n,done_rate", [(1, 1.0), (1023, 0.01), (1024, 0.01), (1025, 0.0), (5000, 0.0), (6144, 0.5), (6145, 1.0),
                                         (12288, 0.0), (18433, 0.002), (50_000, 0.7), (300_000, 1 / 300), (2_000_000, 1 / 3000),
                                         (8_400_003, 1 / 300)])  # the last one is above the dispatcher's 2^23-step switch
def test_gae_flat_matches_oracle_at_scale(E, n, done_rate, entry):
    rng = np.random.default_rng(n)
    r = (rng.integers(0, 64, n) * 4 * (rng.random(n) < 0.4)).astype(np.float32)
    v = (rng.standard_normal(n) * 10).astype(np.float32)
    d = (rng.random(n) < done_rate).astype(np.uint8)
    want_a, want_r = CO.gae(r, v, d, 0.99, 0.95)
    adv, ret, mom = E.gae_flat(dev(r), dev(v), dev(d), 0.99, 0.95, entry=entry)
    np.testing.assert_array_equal(adv.cpu().numpy(), want_a)
    np.testing.assert_array_equal(ret.cpu().numpy(), want_r)
    m = mom.cpu().numpy()
    assert m[0] == n
    np.testing.assert_allclose(m[1], want_a.astype(np.float64).sum(), rtol=1e-9, atol=1e-6)
    np.testing.assert_allclose(m[4], (want_r.astype(np.float64) ** 2).sum(), rtol=1e-9)


@pytest.mark.parametrize("t_steps,n", [(37, 3000), (37, 2944), (1, 128), (16, 128), (17, 1024), (48, 256), (129, 2944)])
def test_gae_time_major_matches_oracle_per_env(E, t_steps, n):
    """n a multiple of 128 takes the bulk-copy ring kernel (stages of 16 rows: step counts below, at and across the stage
    size and the ring depth), anything else one lane per env with register-batched loads; both bit-identical."""
    rng = np.random.default_rng(47 + t_steps)
    r = (rng.integers(0, 16, (t_steps, n)) * 4).astype(np.float32)
    v = rng.standard_normal((t_steps, n)).astype(np.float32)
    d = rng.random((t_steps, n)) < 0.05
    meta = (d.astype(np.uint8) << 6) | rng.integers(0, 64, (t_steps, n)).astype(np.uint8)
    boot = rng.standard_normal(n).astype(np.float32)
    for bootstrap in (None, boot):
        adv, ret, mom = E.gae_time_major(dev(r), dev(v), dev(meta), t_steps, n, None if bootstrap is None else dev(bootstrap), 0.99, 0.95)
        adv, ret = adv.cpu().numpy(), ret.cpu().numpy()
        for e in range(0, n, 97 if n > 200 else 7):
            if bootstrap is None:
                wa, wr = CO.gae(r[:, e], v[:, e], d[:, e], 0.99, 0.95)
            else:  # append the bootstrap state as an extra step and drop it
                wa, wr = O.gae_returns(np.append(r[:, e], 0), np.append(v[:, e], boot[e]), np.append(d[:, e], False), 0.99, 0.95)
                # the appended step has delta = -V and pollutes nothing before it only through last_v / last_gae;
                # rebuild exactly: carry = (0, boot) into step T-1
                last_gae, last_v = np.float32(0), np.float32(boot[e])
                wa = np.zeros(t_steps, np.float32)
                wr = np.zeros(t_steps, np.float32)
                for t in range(t_steps - 1, -1, -1):
                    if d[t, e]:
                        last_gae, last_v = np.float32(0), np.float32(0)
                    delta = np.float32(np.float32(r[t, e] + np.float32(np.float32(0.99) * last_v)) - v[t, e])
                    last_gae = np.float32(delta + np.float32(np.float32(0.99 * 0.95) * last_gae))
                    wa[t] = last_gae
                    wr[t] = np.float32(last_gae + v[t, e])
                    last_v = v[t, e]
            np.testing.assert_array_equal(adv[:, e], wa)
            np.testing.assert_array_equal(ret[:, e], wr)
        assert mom[0].item() == t_steps * n
        np.testing.assert_allclose(mom[1].item(), adv.astype(np.float64).sum(), rtol=1e-9, atol=1e-6)


def test_gae_flat_unaligned_views_and_nonbinary_dones(E):
    rng = np.random.default_rng(53)
    n = 40_000
    r = (rng.integers(0, 64, n + 3) * 4).astype(np.float32)
    v = rng.standard_normal(n + 3).astype(np.float32)
    d = ((rng.random(n + 3) < 0.01) * rng.integers(1, 256, n + 3)).astype(np.uint8)  # any non-zero byte is a done
    want_a, want_r = CO.gae(r[3:], v[3:], d[3:], 0.99, 0.95)
    # 4-byte aligned but not 16-byte aligned views
    adv, ret, _ = E.gae_flat(dev(r)[3:], dev(v)[3:], dev(d)[3:], 0.99, 0.95)
    np.testing.assert_array_equal(adv.cpu().numpy(), want_a)
    np.testing.assert_array_equal(ret.cpu().numpy(), want_r)
    adv, ret, _ = E.gae_flat(dev(r[3:].copy()), dev(v[3:].copy()), dev(d[3:].copy()), 0.99, 0.95)
    np.testing.assert_array_equal(adv.cpu().numpy(), want_a)


def test_gae_host_entry_point(E, golden_ppo):
    g = golden_ppo
    adv, ret = E.gae_host(g["gae_default_rewards"], g["gae_default_values"], g["gae_default_dones"], 0.99, 0.95, False)
    np.testing.assert_array_equal(adv, g["gae_default_adv"])
    adv, ret = E.gae_host(g["gae_default_rewards"], g["gae_default_values"], g["gae_default_dones"], 0.99, 0.95, True)
    np.testing.assert_allclose(adv, g["gae_default_adv_norm"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(ret, g["gae_default_ret_norm"], rtol=1e-5, atol=1e-6)


def test_row_moments_match_reference_running_stats(E, golden_ppo):
    g = golden_ppo
    x = g["rs_push2"]
    out = E.row_moments(dev(x)).cpu().numpy()
    np.testing.assert_allclose(out[:, 0], x.shape[1])
    np.testing.assert_allclose(out[:, 1], x.mean(axis=1), rtol=1e-12)
    np.testing.assert_allclose(out[:, 2], x.var(axis=1), rtol=1e-10)


def test_no_cpu_fallback_errors_are_loud(E):
    with pytest.raises(RuntimeError):
        E.play(0, torch.zeros((2, 2), dtype=torch.int32, device="cuda"), 4, 0, 4, 1)  # n_subs < 3
    with pytest.raises(RuntimeError):
        E.env_init(u32([1, 2]), 4, 3, 2, 1)  # env_lo + n > batch


@pytest.mark.parametrize("n,first,m", [(1, 0, 1), (5, 0, 5), (17, 3, 14), (4097, 0, 4097), (100003, 777, 50001),
                                         (31_000_000, 0, 300_000), ((1 << 40) + 12345, 1 << 39, 70001)])
def test_random_subset_matches_oracle(E, n, first, m):
    """g2048_random_subset vs the oracle's restatement, bit-exact; the outputs are distinct positions of [0, n)."""
    key = (0x9E3779B9, n & 0xFFFFFFFF)
    got = E.random_subset(n, m, key, "cuda", first=first).cpu().numpy()
    np.testing.assert_array_equal(got, O.random_subset(key, n, first, m))
    assert got.min() >= 0 and got.max() < n and len(np.unique(got)) == m
    if first == 0 and m == n:
        assert np.array_equal(np.sort(got), np.arange(n))


def test_random_subset_rejects_bad_ranges(E):
    for n, first, m in ((10, 5, 6), (-1, 0, 0), (10, -1, 2)):
        with pytest.raises(RuntimeError):
            E.random_subset(n, m, 1, "cuda", first=first)


@pytest.mark.parametrize("row_shape,dtype", [((), torch.uint8), ((3,), torch.uint8), ((1,), torch.float32), ((5, 4), torch.float32),
                                             ((16, 31), torch.float32), ((7,), torch.int16)])
def test_compact_rows_matches_store_batch_indices(E, row_shape, dtype):
    """g2048_first_done_rows + g2048_compact_rows (byte, 4-byte and 16-byte copy widths) vs the oracle's store_batch."""
    rng = np.random.default_rng(len(row_shape) + 17)
    for n_envs, t_steps in ((1, 1), (5, 33), (130, 64), (257, 100)):
        term = rng.random((n_envs, t_steps)) < 0.03
        term[0] = False
        src = torch.from_numpy(rng.integers(0, 200, (n_envs, t_steps, *row_shape))).to(dtype).cuda()
        lengths = E.first_done_rows(torch.from_numpy(term).cuda())
        e, st = O.store_batch_indices(term)
        want_len = np.array([np.flatnonzero(r)[0] + 1 if r.any() else 0 for r in term])
        np.testing.assert_array_equal(lengths.cpu().numpy(), want_len)
        offsets = E.exclusive_scan(lengths)
        assert int(offsets[-1]) == len(e)
        out = E.compact_rows(src, lengths, offsets, len(e))
        np.testing.assert_array_equal(out.cpu().numpy(), src.cpu().numpy()[e, st])
