"""Pins the oracle (test infrastructure) to the reference's golden artefacts.  CPU only."""
from collections import Counter

import numpy as np
import pytest

from oracle import c_oracle as CO
from oracle import pgx2048_oracle as O


# ----------------------------------------------------------------------------- RNG KATs
@pytest.mark.parametrize(
    "k,x,want",
    [
        ((0, 0), (0, 0), (0x6B200159, 0x99BA4EFE)),
        ((0xFFFFFFFF, 0xFFFFFFFF), (0xFFFFFFFF, 0xFFFFFFFF), (0x1CB996FC, 0xBB002BE7)),
        ((0x13198A2E, 0x03707344), (0x243F6A88, 0x85A308D3), (0xC4923A9C, 0x483DF7A0)),
    ],
)
def test_threefry_random123_kats(k, x, want):
    y = O.threefry2x32(k[0], k[1], x[0], x[1])
    assert (int(y[0]), int(y[1])) == want
    assert CO.threefry2x32(k[0], k[1], x[0], x[1]) == want


def test_jax_random_kats_original_mode():
    k = (np.array([0], np.uint32), np.array([0], np.uint32))
    s = O.split(k, 2, O.ORIGINAL)
    # the values printed in the JAX documentation for split(key(0))
    assert [int(s[0][0, 0]), int(s[1][0, 0])] == [4146024105, 967050713]
    assert [int(s[0][0, 1]), int(s[1][0, 1])] == [2718843009, 1272950319]
    assert O.uniform(k, 1, O.ORIGINAL)[0, 0] == np.float32(0.41845703)
    np.testing.assert_array_equal(
        O.uniform(k, 4, O.ORIGINAL)[0], np.array([0.9653214, 0.22515893, 0.63302994, 0.29638183], np.float32)
    )
    np.testing.assert_array_equal(CO.split([0, 0], 2, CO.ORIGINAL), [[4146024105, 967050713], [2718843009, 1272950319]])


def test_jax_random_kats_partitionable_mode():
    k = (np.array([0], np.uint32), np.array([0], np.uint32))
    s = O.split(k, 2, O.PARTITIONABLE)
    assert [int(s[0][0, 0]), int(s[1][0, 0])] == [0x6B200159, 0x99BA4EFE]
    assert [int(s[0][0, 1]), int(s[1][0, 1])] == [0x375F238F, 0xCDDB151D]
    np.testing.assert_array_equal(
        O.uniform(k, 4, O.PARTITIONABLE)[0], np.array([0.947667, 0.9785799, 0.33229148, 0.46866846], np.float32)
    )
    k42 = (np.array([0], np.uint32), np.array([42], np.uint32))
    s = O.split(k42, 2, O.PARTITIONABLE)
    assert [int(s[0][0, 0]), int(s[1][0, 0]), int(s[0][0, 1]), int(s[1][0, 1])] == [0x6D3E048F, 0x1022172D, 0x03D7B32D, 0xADD083F4]


@pytest.mark.parametrize("mode", [O.ORIGINAL, O.PARTITIONABLE])
@pytest.mark.parametrize("n", [1, 2, 3, 4, 7, 100])
def test_split_c_matches_numpy(mode, n):
    k = (np.array([0xDEADBEEF], np.uint32), np.array([0x12345678], np.uint32))
    s = O.split(k, n, mode)
    want = np.stack([s[0][0], s[1][0]], axis=1)
    np.testing.assert_array_equal(CO.split([0xDEADBEEF, 0x12345678], n, mode), want)


# ----------------------------------------------------------------------------- golden trajectories
@pytest.mark.parametrize("policy,name", [("drul", "drul_boards"), ("random", "random_boards")])
def test_numpy_oracle_reproduces_svg(golden_svg, golden_hist, policy, name):
    out = O.rollout(O.KeyChain(0, O.ORIGINAL), 4, policy)
    boards = np.stack([s.board for s in out["states"]])
    np.testing.assert_array_equal(boards, golden_svg[name])
    length, score, max_tile = O.episode_summary(out)
    want = golden_hist["svg_seed0_batch4_original_mode"][policy]
    assert length.tolist() == want["lengths"]
    assert score.tolist() == want["scores"]
    assert max_tile.tolist() == want["max_tiles"]
    assert [int(a[0]) for a in out["actions"][:12]] == want["first_actions_env0"]


def test_known_init_state_original_mode():
    st = O.env_init(O.KeyChain(0, O.ORIGINAL).next_batch_keys(4), O.ORIGINAL)
    want = [{8: 1, 12: 1}, {1: 1, 14: 1}, {8: 1, 10: 2}, {8: 1, 12: 1}]
    for e in range(4):
        assert {int(i): int(st.board[e, i]) for i in np.nonzero(st.board[e])[0]} == want[e]
    np.testing.assert_array_equal(st.legal_action_mask, [[0, 1, 1, 1], [1, 1, 1, 1], [1, 1, 1, 1], [0, 1, 1, 1]])


@pytest.mark.parametrize("policy,cpol,name", [("drul", CO.DRUL, "drul_boards"), ("random", CO.RANDOM, "random_boards")])
def test_c_oracle_reproduces_svg(golden_svg, golden_hist, policy, cpol, name):
    r = CO.play(0, 4, cpol, CO.ORIGINAL, first_actions=True)
    want = golden_hist["svg_seed0_batch4_original_mode"][policy]
    assert r["lengths"].tolist() == want["lengths"]
    assert r["scores"].tolist() == want["scores"]
    assert r["longest"] == want["loop_steps"]
    assert r["first_actions"][0, :12].tolist() == want["first_actions_env0"]
    g = golden_svg[name]
    for e in range(4):  # the SVG freezes a finished env, so its last frame is each env's final board
        np.testing.assert_array_equal(r["final_boards"][e], g[-1, e])


@pytest.mark.parametrize("policy,cpol", [("random", CO.RANDOM), ("drul", CO.DRUL)])
def test_c_oracle_reproduces_png_histograms(golden_hist, policy, cpol):
    c = Counter()
    for seed in golden_hist["seeds"]:
        r = CO.play(seed, golden_hist["batch_size"], cpol, CO.PARTITIONABLE)
        c.update((1 << r["final_boards"].max(axis=1).astype(np.int64)).tolist())
    got = {str(k): round(v / 10, 1) for k, v in sorted(c.items())}
    assert got == golden_hist[policy]


def test_numpy_oracle_reproduces_png_histogram_random_first_seeds(golden_hist):
    # the numpy path is slow; two seeds against the C path (which is pinned to all ten above)
    for seed in golden_hist["seeds"][:2]:
        out = O.rollout(O.KeyChain(seed, O.PARTITIONABLE), 100, "random")
        length, score, _ = O.episode_summary(out)
        r = CO.play(seed, 100, CO.RANDOM, CO.PARTITIONABLE)
        np.testing.assert_array_equal(length, r["lengths"])
        np.testing.assert_array_equal(score, r["scores"])
        np.testing.assert_array_equal(out["states"][-1].board, r["final_boards"])


# ----------------------------------------------------------------------------- step semantics, C vs numpy
@pytest.mark.parametrize("mode", [O.ORIGINAL, O.PARTITIONABLE])
def test_step_c_matches_numpy_on_random_boards(mode):
    rng = np.random.default_rng(5)
    n = 4000
    boards = rng.integers(0, 8, (n, 16)) * (rng.random((n, 16)) < 0.7)
    boards[:50] = rng.integers(1, 12, (50, 16))  # full boards: terminal / illegal paths
    masks = O.exact_legal(boards)
    done = ~masks.any(axis=1)
    masks = np.where(done[:, None], True, masks)
    actions = rng.integers(0, 4, n)  # includes illegal actions
    keys = rng.integers(0, 2**32, (n, 2), dtype=np.uint64).astype(np.uint32)
    st = O.State(boards.astype(np.int32), masks, np.zeros((n, 1), np.float32), done, np.zeros(n, bool))
    want = O.env_step(st, actions, (keys[:, 0], keys[:, 1]), mode)
    b, m, d, r = CO.env_step(boards, masks, done, actions, keys, mode)
    np.testing.assert_array_equal(b, want.board)
    np.testing.assert_array_equal(m.astype(bool), want.legal_action_mask)
    np.testing.assert_array_equal(d.astype(bool), want.terminated)
    np.testing.assert_array_equal(r, want.rewards[:, 0])
    assert (r == -1).any() and d.any() and (r > 0).any()


def test_step_semantics_hand_built():
    # merge once per pair, left to right: [2,2,2,2] -> [4,4,0,0], reward 8
    b = np.zeros((1, 16), np.int32)
    b[0, :4] = 1
    moved, rew = O.move(b, [0])
    assert moved[0, :4].tolist() == [2, 2, 0, 0] and rew[0] == 8
    # [2,2,4,0] right -> [0,0,4,4]? no: merge toward the move direction: [0,0,4,4] with reward 4
    b[0, :4] = [1, 1, 2, 0]
    moved, rew = O.move(b, [2])
    assert moved[0, :4].tolist() == [0, 0, 2, 2] and rew[0] == 4
    # up / down on a column
    b[:] = 0
    b[0, [0, 4, 8, 12]] = [3, 3, 0, 3]
    moved, rew = O.move(b, [1])
    assert moved[0, [0, 4, 8, 12]].tolist() == [4, 3, 0, 0] and rew[0] == 16
    moved, rew = O.move(b, [3])
    assert moved[0, [0, 4, 8, 12]].tolist() == [0, 0, 3, 4] and rew[0] == 16
    # frozen env: zero reward, unchanged board
    st = O.State(b.copy(), np.ones((1, 4), bool), np.full((1, 1), 5, np.float32), np.ones(1, bool), np.zeros(1, bool))
    nxt = O.env_step_given(st, [0], np.float32([0.3]), np.float32([0.3]))
    np.testing.assert_array_equal(nxt.board, b)
    assert nxt.rewards[0, 0] == 0 and nxt.terminated[0]


def test_spawn_rule_value_threshold_and_position():
    b = np.zeros((3, 16), np.int32)
    b[:, [0, 5]] = 3
    # (1-u) > 0.9f  <=>  a 4-tile (exponent 2)
    u_val = np.float32([0.05, 0.2, np.float32(1.0) - np.float32(0.9)])
    u_pos = np.float32([0.0, 0.999999, 0.5])
    out = O.add_random_given(b, u_pos, u_val)
    new = [(int(np.nonzero(out[i] != b[i])[0][0]), int(out[i][out[i] != b[i]][0])) for i in range(3)]
    # 14 empty cells; r = 14*(1-u): u=0 -> 14th empty cell (index 15); u~1 -> 1st empty (index 1); u=.5 -> 7th (index 8)
    assert new[0] == (15, 2)
    assert new[1] == (1, 1)
    assert new[2][0] == 8


# ----------------------------------------------------------------------------- policies
def test_act_kats_from_reference_tests():
    # tests/actions/test_act_drul.py:40-48 and test_act_randomly.py:42-53
    mask = np.array([[True, False, True, False]])
    assert O.act_drul(mask)[0] == 2
    for s in range(20):
        k = (np.array([0], np.uint32), np.array([s], np.uint32))
        for mode in (O.ORIGINAL, O.PARTITIONABLE):
            a, lp = O.act_randomly(k, mask, mode)
            assert a[0] in (0, 2) and lp[0] == np.log(np.float32(0.5))
            a2, _ = O.act_randomly(k, mask, mode, shortcut=True)
            assert a2[0] == a[0]
            a3, lp3 = CO.act(np.array([[0, s]]), mask, CO.RANDOM, mode)
            assert a3[0] == a[0] and lp3[0] == lp[0]
    assert O.act_drul(np.zeros((1, 4), bool))[0] == 3


def test_act_random_shortcut_and_c_agree_at_scale():
    rng = np.random.default_rng(9)
    n = 20000
    masks = rng.random((n, 4)) < 0.6
    masks[masks.sum(1) == 0, 2] = True
    keys = rng.integers(0, 2**32, (n, 2), dtype=np.uint64).astype(np.uint32)
    for mode in (O.ORIGINAL, O.PARTITIONABLE):
        a, lp = O.act_randomly((keys[:, 0], keys[:, 1]), masks, mode)
        a2, _ = O.act_randomly((keys[:, 0], keys[:, 1]), masks, mode, shortcut=True)
        a3, lp3 = CO.act(keys, masks, CO.RANDOM, mode)
        np.testing.assert_array_equal(a, a2)
        np.testing.assert_array_equal(a, a3)
        np.testing.assert_array_equal(lp, lp3)
        assert masks[np.arange(n), a].all()


# ----------------------------------------------------------------------------- GAE / buffer / stats vs the reference
TAGS = ["default", "short_eps", "undiscounted", "lowlam", "open_tail", "two"]


@pytest.mark.parametrize("tag", TAGS)
def test_gae_oracles_match_reference_bit_exact(golden_ppo, tag):
    g = golden_ppo
    gamma, lam = g[f"gae_{tag}_params"]
    args = (g[f"gae_{tag}_rewards"], g[f"gae_{tag}_values"], g[f"gae_{tag}_dones"], gamma, lam)
    for fn in (O.gae_returns, CO.gae):
        adv, ret = fn(*args)
        np.testing.assert_array_equal(adv, g[f"gae_{tag}_adv"])
        np.testing.assert_array_equal(ret, g[f"gae_{tag}_ret"])
    np.testing.assert_allclose(O.normalize(adv), g[f"gae_{tag}_adv_norm"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(O.normalize(ret), g[f"gae_{tag}_ret_norm"], rtol=1e-5, atol=1e-6)


def test_store_batch_indices_match_reference(golden_ppo):
    g = golden_ppo
    e1, s1 = O.store_batch_indices(g["rb_terminations"])
    e2, s2 = O.store_batch_indices(g["rb_terminations"][:3])
    e, s = np.concatenate([e1, e2]), np.concatenate([s1, s2])
    assert len(e) == int(g["rb_size"][0])
    np.testing.assert_array_equal(g["rb_rewards"][e, s], g["rb_out_rewards"])
    np.testing.assert_array_equal(g["rb_terminations"][e, s], g["rb_out_terminations"])
    obs = O.observe(g["rb_boards"][e, s]).reshape(-1, 16, 31).astype(np.float32)
    np.testing.assert_array_equal(obs, g["rb_out_observations"])


def test_running_stats_merge_matches_reference(golden_ppo):
    g = golden_ppo
    n = np.zeros((3, 1))
    mean = np.zeros((3, 1))
    var = np.zeros((3, 1))
    for i in range(4):
        x = g[f"rs_push{i}"]
        n, mean, var = O.running_stats_merge(n, mean, var, x.shape[1], x.mean(1, keepdims=True), x.var(1, keepdims=True))
        np.testing.assert_allclose(mean, g[f"rs_mean{i}"], rtol=1e-12)
        np.testing.assert_allclose(var, g[f"rs_var{i}"], rtol=1e-12)
        np.testing.assert_array_equal(n, g[f"rs_n{i}"])


def test_logprob_formula_matches_reference(golden_ppo):
    g = golden_ppo
    for pre in ("lp", "lp2"):
        masked = O.mask_logits(g[f"{pre}_raw_logits"], g["lp_masks"])
        np.testing.assert_array_equal(masked, g[f"{pre}_masked_logits"])
        lg = np.maximum(masked, np.finfo(np.float32).min)
        m = lg.max(-1, keepdims=True)
        lse = m + np.log(np.exp(lg - m).sum(-1, keepdims=True))
        lp = (lg - lse)[np.arange(lg.shape[0]), g[f"{pre}_actions"]]
        np.testing.assert_allclose(lp, g[f"{pre}_log_probs"], rtol=1e-5, atol=1e-6)


def test_random_subset_is_a_keyed_bijection():
    """oracle.random_subset (the restatement of g2048_random_subset): every n gives a permutation of [0, n), windows
    of it are its slices, different keys give different permutations, and a subset covers the range evenly."""
    for n in (1, 2, 3, 4, 5, 16, 17, 255, 256, 257, 4097, 100003):
        full = O.random_subset((11, 22), n, 0, n)
        assert np.array_equal(np.sort(full), np.arange(n)), n
        if n > 20:
            np.testing.assert_array_equal(O.random_subset((11, 22), n, 7, 9), full[7:16])
            assert not np.array_equal(full, O.random_subset((11, 23), n, 0, n))
    n, m = 31_000_000, 300_000  # C4's buffer and configs/trainer/default.yaml max_samples_per_epoch
    sub = O.random_subset((0, 2048), n, 0, m)
    assert len(np.unique(sub)) == m and sub.min() >= 0 and sub.max() < n
    hist = np.bincount(sub * 100 // n, minlength=100)  # 3 000 expected per bin, sigma ~ 55
    assert np.abs(hist - m / 100).max() < 6 * np.sqrt(m / 100)
    # no visible order: consecutive outputs are uncorrelated
    c = np.corrcoef(sub[:-1].astype(np.float64), sub[1:].astype(np.float64))[0, 1]
    assert abs(c) < 0.01
