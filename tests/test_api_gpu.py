"""The reference-facing Python API on the GPU: mirrors the reference's own tests
(tests/actions, tests/runs, tests/ppo/test_rollout_buffer.py, tests/ppo/test_data_loader.py,
tests/test_running_stats_vec.py, tests/integration) and adds parity against the oracle."""
from collections import Counter

import numpy as np
import pytest
import torch

from oracle import c_oracle as CO
from oracle import pgx2048_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def G():
    import g2048

    return g2048


def mock_act_fn(keys, obs, mask):
    """The reference tests' stand-in policy: a random action that IGNORES the mask
    (tests/runs/test_batch_runner.py:12-16), here on torch tensors."""
    n = mask.shape[0]
    words = keys.to(torch.int64) & 0xFFFFFFFF
    action = ((words[:, 0] ^ words[:, 1]) % 4).to(torch.int32)
    return action, torch.zeros(n, device=mask.device), torch.zeros(n, device=mask.device)


# ------------------------------------------------------------------------------------- src.* drop-in imports
def test_reference_import_paths():
    from src.actions import act_drul, act_randomly  # noqa: F401
    from src.env_definitions import ACTION_DIM, BOARD_DIM, BOARD_FLAT_DIM, OBS_DIM
    from src.ppo.data_loader import PPODataset, create_ppo_dataloader  # noqa: F401
    from src.ppo.rollout_buffer import RolloutBuffer  # noqa: F401
    from src.ppo.torch_action_wrapper import TorchActionFunction  # noqa: F401
    from src.runs import BatchRunner, run_actions_batch, run_actions_max_tile  # noqa: F401
    from src.stats import RunningStatsVec  # noqa: F401

    assert (OBS_DIM, BOARD_DIM, BOARD_FLAT_DIM, ACTION_DIM) == (31, (4, 4), 16, 4)


# ------------------------------------------------------------------------------------- actions (tests/actions/*)
@pytest.mark.parametrize("batch", [1, 10, 100])
def test_act_functions_shapes_and_ranges(G, batch):
    keys = G.keys.split(G.keys.key(0), batch)
    obs = torch.zeros((batch, 4, 4, 31), dtype=torch.bool)
    mask = torch.ones((batch, 4), dtype=torch.bool)
    for fn in (G.act_randomly, G.act_drul):
        action, log_prob, value = fn(keys, obs, mask)
        assert action.shape == (batch,) and value is None
        assert ((action >= 0) & (action < 4)).all()
    a, lp, v = G.act_randomly(keys[0], obs[0], mask[0])
    assert a.shape == () and lp.shape == () and v is None


def test_act_kats(G):
    mask = np.array([True, False, True, False])
    obs = np.zeros((4, 4, 31), bool)
    for seed in range(10):
        a, lp, _ = G.act_randomly(G.keys.key(seed), obs, mask)
        assert a.item() in (0, 2) and np.isclose(lp.item(), np.log(0.5))
    a, lp, v = G.act_drul(G.keys.key(0), obs, mask)
    assert a.item() == 2 and lp is None and v is None
    with pytest.raises(AssertionError):
        G.act_drul(G.keys.key(0), np.zeros((4, 4, 30), bool), mask)


# ------------------------------------------------------------------------------------- BatchRunner (tests/runs/*)
def test_batch_runner_act_fn_property_and_errors(G):
    runner = G.BatchRunner(init_seed=0)
    assert runner.act_fn is None
    with pytest.raises(ValueError, match="The action function is not set"):
        runner.run_actions_batch(4)
    with pytest.raises(ValueError, match="The action function is not set"):
        runner.run_rollout_batch(4)
    runner.act_fn = mock_act_fn
    assert runner.act_fn is not None
    with pytest.raises((ValueError, RuntimeError)):
        runner.run_actions_batch(0)


@pytest.mark.parametrize("batch", [1, 10, 100])
def test_run_actions_batch_shapes_with_mask_ignoring_policy(G, batch):
    runner = G.BatchRunner(init_seed=0, act_fn=mock_act_fn)
    obs, actions, masks, log_probs, values, rewards, terms = runner.run_actions_batch(batch)
    t = obs.shape[1]
    assert obs.shape == (batch, t, 4, 4, 31) and obs.dtype == bool
    assert actions.shape == (batch, t) and masks.shape == (batch, t, 4)
    assert log_probs.shape == values.shape == rewards.shape == terms.shape == (batch, t)
    assert terms[:, -1].all()
    assert np.isfinite(log_probs).all() and np.isfinite(values).all()
    assert set(np.unique(masks)) <= {0, 1} and set(np.unique(terms)) <= {0, 1}
    # the mask-ignoring policy runs into Pgx's illegal-action rule: reward -1 ends the episode
    assert (rewards == -1).any()
    first = terms.argmax(axis=1)
    ended_illegal = rewards[np.arange(batch), first] == -1
    taken = actions[np.arange(batch), first]
    assert (~masks[np.arange(batch), first, taken][ended_illegal]).all()
    assert (obs.reshape(batch, t, 16, 31).sum(-1) == 1).all()


def test_run_actions_batch_is_deterministic(G):
    outs = []
    for _ in range(2):
        runner = G.BatchRunner(init_seed=42, act_fn=G.act_randomly)
        outs.append(runner.run_actions_batch(16))
    for a, b in zip(*outs):
        if a is None:
            assert b is None
        else:
            np.testing.assert_array_equal(a, b)
    # the chain advances: a second call on the same runner gives a different batch
    again = runner.run_actions_batch(16)
    assert again[0].shape != outs[0][0].shape or not np.array_equal(again[0], outs[0][0])


@pytest.mark.parametrize("mode", ["original", "partitionable"])
@pytest.mark.parametrize("policy", ["random", "drul"])
def test_run_actions_batch_matches_numpy_oracle(G, mode, policy):
    imode = 0 if mode == "original" else 1
    fn = G.act_randomly if policy == "random" else G.act_drul
    runner = G.BatchRunner(init_seed=3, act_fn=fn, rng_mode=mode)
    ref_chain = O.KeyChain(3, imode)
    for _ in range(2):  # two consecutive runs: the chain position carries over like the reference's self.key
        obs, actions, masks, log_probs, values, rewards, terms = runner.run_actions_batch(32)
        ref = O.rollout(ref_chain, 32, policy)
        t = len(ref["states"])
        assert obs.shape[1] == t
        np.testing.assert_array_equal(obs, np.stack([O.observe(b) for b in ref["boards"]], axis=1))
        np.testing.assert_array_equal(actions, np.stack(ref["actions"], axis=1))
        np.testing.assert_array_equal(masks, np.stack(ref["masks"], axis=1))
        np.testing.assert_array_equal(rewards, np.concatenate([s.rewards for s in ref["states"]], axis=1))
        np.testing.assert_array_equal(terms, np.stack([s.terminated for s in ref["states"]], axis=1))
        assert values is None
        if policy == "random":
            np.testing.assert_allclose(log_probs, np.stack(ref["log_probs"], axis=1), rtol=1e-6)
        else:
            assert log_probs is None
        np.testing.assert_array_equal(runner.key, np.array(ref_chain.key, np.uint32))


def test_runner_state_dict_resumes_the_key_chain(G):
    a = G.BatchRunner(init_seed=9, act_fn=G.act_randomly)
    a.run_packed_batch(16)
    saved = a.state_dict()
    want = a.run_packed_batch(16)
    b = G.BatchRunner(init_seed=0, act_fn=G.act_randomly)
    b.load_state_dict(saved)
    got = b.run_packed_batch(16)
    assert torch.equal(got.boards, want.boards) and torch.equal(got.meta, want.meta)
    assert b.state_dict() == a.state_dict() and saved["position"] > 0


def test_golden_svg_through_the_public_api(G, golden_svg):
    """run_actions_batch(0, 4, act_drul / act_randomly) == the frames of the reference's SVGs."""
    for fn, name in ((G.act_drul, "drul_boards"), (G.act_randomly, "random_boards")):
        states = G.run_actions_batch(0, 4, fn, rng_mode="original")
        boards = np.stack([G.engine.boards_numpy(s.boards) for s in states])
        np.testing.assert_array_equal(boards, golden_svg[name])
        assert all(isinstance(s, G.State) for s in states)
        assert states[0].observation.shape == (4, 4, 4, 31) and states[0].legal_action_mask.shape == (4, 4)
        assert states[0].rewards.shape == (4, 1) and states[-1].terminated.all() and not states[-1].truncated.any()


def test_run_rollout_batch_includes_init_state_and_png_histogram(G, golden_hist):
    c = Counter()
    for seed in golden_hist["seeds"][:3]:
        runner = G.BatchRunner(seed, G.act_drul)
        states = runner.run_rollout_batch(100)
        assert not states[0].terminated.any() and (states[0].rewards == 0).all()
        final = G.engine.boards_numpy(states[-1].boards)
        c.update((1 << final.max(axis=1).astype(np.int64)).tolist())
        ref = CO.play(seed, 100, CO.DRUL, CO.PARTITIONABLE)
        np.testing.assert_array_equal(final, ref["final_boards"])
        assert len(states) == ref["longest"] + 1


def test_run_stats_batch_and_max_tile(G, golden_hist):
    # the protocol behind the README histograms, through run_stats_batch (persistent kernel)
    hist = Counter()
    for seed in golden_hist["seeds"]:
        out = G.BatchRunner(seed, G.act_randomly).run_stats_batch(100)
        for k, v in out["summary"]["max_tile_hist"].items():
            hist[k] += v
    assert {str(k): round(v / 10, 1) for k, v in sorted(hist.items())} == golden_hist["random"]
    # run_actions_max_tile: sample count as tests/runs/test_run_actions_max_tile.py:18-29
    stats = G.run_actions_max_tile(0, 10, 35, G.act_drul)
    assert stats.num_samples[0] == 30
    exact = G.run_actions_max_tile(7, 64, 128, G.act_randomly, exact_reference_quirk=False)
    tiles = []
    key = None
    for _ in range(2):
        r = CO.play(7, 64, CO.RANDOM, CO.PARTITIONABLE, key=key)
        tiles.append(1 << r["final_boards"].max(axis=1).astype(np.int64))
        key, _ = CO.chain(key if key is not None else [0, 7], CO.PARTITIONABLE, 1 + 2 * r["longest"])
    tiles = np.concatenate(tiles)
    np.testing.assert_allclose(exact.mean[0, 0], tiles.mean(), rtol=1e-12)
    np.testing.assert_allclose(exact.variance[0, 0], tiles.var(), rtol=1e-10)
    # with the reference's quirk the longest-lived env misses its last merge: max tile can only be <=
    quirk = G.run_actions_max_tile(7, 64, 128, G.act_randomly)
    assert quirk.num_samples[0] == 128 and quirk.mean[0, 0] <= exact.mean[0, 0]


def test_batch_runner_shards_reproduce_the_full_batch(G):
    full = G.BatchRunner(11, G.act_randomly).run_packed_batch(64)
    parts = [G.BatchRunner(11, G.act_randomly, shard=(r, 4)).run_packed_batch(64) for r in range(4)]
    t = full.t_steps
    # single process: each shard stops on its own longest episode; compare the common prefix per env
    for r, p in enumerate(parts):
        lo = 16 * r
        tt = min(t, p.t_steps)
        assert torch.equal(full.boards[:tt, lo : lo + 16], p.boards[:tt])
        assert torch.equal(full.final_boards[lo : lo + 16], p.final_boards)


# ------------------------------------------------------------------------------------- RunningStatsVec
def test_running_stats_vec_matches_numpy(G):
    rng = np.random.default_rng(0)
    data = rng.standard_normal((100, 10000))
    for pushes in range(1, 6):
        rs = G.RunningStatsVec()
        for part in np.array_split(data, pushes, axis=1):
            rs.push(part)
        assert np.allclose(rs.mean[:, 0], data.mean(axis=1)) and np.allclose(rs.variance[:, 0], data.var(axis=1))
        assert np.allclose(rs.std[:, 0], data.std(axis=1)) and (rs.num_samples == 10000).all()
    rs.clear()
    assert rs.num_samples.sum() == 0 and rs.mean == 0.0 and rs.variance == 0.0
    with pytest.raises(ValueError, match="2 dimensions"):
        rs.push(np.zeros(5))


def test_running_stats_vec_matches_reference_fixture(G, golden_ppo):
    rs = G.RunningStatsVec()
    for i in range(4):
        rs.push(golden_ppo[f"rs_push{i}"])
        np.testing.assert_allclose(rs.mean, golden_ppo[f"rs_mean{i}"], rtol=1e-12)
        np.testing.assert_allclose(rs.variance, golden_ppo[f"rs_var{i}"], rtol=1e-10)
        np.testing.assert_array_equal(rs.num_samples, golden_ppo[f"rs_n{i}"])


# ------------------------------------------------------------------------------------- RolloutBuffer
def test_rollout_buffer_matches_reference_fixture(G, golden_ppo):
    g = golden_ppo
    obs = O.observe(g["rb_boards"].reshape(-1, 16)).reshape(*g["rb_boards"].shape[:2], 4, 4, 31)
    onehot = np.eye(4, dtype=np.float32)[g["rb_actions"]]
    rb = G.RolloutBuffer(observation_dim=31, observation_length=16, action_dim=4)
    args = (obs, onehot, g["rb_masks"], g["rb_rewards"], g["rb_values"], g["rb_log_probs"], g["rb_terminations"])
    rb.store_batch(*args)
    assert rb.buffer_size == 39
    rb.store_batch(*(a[:3] for a in args))
    assert rb.buffer_size == int(g["rb_size"][0])
    data = rb.get_buffer_data()
    for k in ("observations", "actions", "action_masks", "rewards", "values", "log_probs", "terminations"):
        assert data[k].dtype == g[f"rb_out_{k}"].dtype and data[k].shape == g[f"rb_out_{k}"].shape, k
        np.testing.assert_array_equal(data[k], g[f"rb_out_{k}"], err_msg=k)
    rb.reset()
    assert rb.buffer_size == 0 and rb.get_buffer_data()["rewards"].shape == (0,)


def test_rollout_buffer_semantics_of_the_reference_tests(G):
    rb = G.RolloutBuffer(31, 16, 4)
    b, t = 2, 6
    obs = np.zeros((b, t, 16, 31), np.float32)
    obs[..., 0] = 1
    z = np.zeros((b, t), np.float32)
    term = np.zeros((b, t), bool)
    rb.store_batch(obs, np.zeros((b, t, 4), np.float32), np.ones((b, t, 4), bool), z, z, z, term)
    assert rb.buffer_size == 0  # no termination -> nothing stored (test_rollout_buffer.py:129-130)
    term[0, 3] = True
    term[1, 4] = True
    rewards = np.arange(b * t, dtype=np.float32).reshape(b, t)
    rb.store_batch(obs.reshape(b, t, 4, 4, 31), np.zeros((b, t, 4), np.float32), np.ones((b, t, 4), bool), rewards, z, z, term)
    assert rb.buffer_size == 9  # 4 + 5 (:91-92)
    data = rb.get_buffer_data()
    np.testing.assert_array_equal(data["rewards"], [0, 1, 2, 3, 6, 7, 8, 9, 10])  # env-major order (:706-717)
    assert data["observations"].dtype == np.float32 and data["terminations"].dtype == bool
    with pytest.raises(ValueError, match="at least 2 dimensions"):
        rb.store_batch(np.zeros(5), z, z, z, z, z, term)
    with pytest.raises(ValueError, match="Failed to reshape observations"):
        rb.store_batch(np.zeros((b, t, 7, 31)), z, z, z, z, z, term)


def test_rollout_buffer_is_generic_in_observation_and_action_shapes(G):
    """The reference buffer stores any (B, T, ...) rows (its tests use 4- and 10-dimensional observations,
    test_rollout_buffer.py:60-125): the generic device path against the oracle's restatement of store_batch."""
    rng = np.random.default_rng(5)
    for obs_len, obs_dim, act_dim, b, t in ((5, 4, 2, 7, 9), ((2, 3), 10, 3, 33, 70), (16, 31, 4, 4, 40), (1, 1, 1, 3, 1)):
        rb = G.RolloutBuffer(obs_dim, obs_len, act_dim)
        for name in ("observation", "action", "action_mask", "reward", "value", "log_prob", "termination"):
            assert getattr(rb, name + "_buffer") == []  # the reference's list attributes exist (and reset() clears them)
        dims = (*obs_len, obs_dim) if isinstance(obs_len, tuple) else (obs_len, obs_dim)
        want = {k: [] for k in ("observations", "actions", "action_masks", "rewards", "values", "log_probs", "terminations")}
        for call in range(3):
            obs = rng.standard_normal((b, t, *dims)).astype(np.float32)
            act = rng.random((b, t, act_dim)).astype(np.float32)
            masks = rng.random((b, t, act_dim)) < 0.5
            r, v, lp = (rng.standard_normal((b, t)).astype(np.float32) for _ in range(3))
            term = rng.random((b, t)) < 0.08
            term[0] = False          # an env that never terminates stores nothing
            if b > 1:
                term[1, 0] = True    # termination at the first step: one step kept
            flat_obs = obs.reshape(b, t, -1) if call == 1 else obs  # flattened input is reshaped (rollout_buffer.py:99-110)
            kept = rb.store_batch(flat_obs, act, masks, r, v, lp, term)
            e, st = O.store_batch_indices(term)
            assert kept == len(e)
            for k, a in (("observations", obs), ("actions", act), ("action_masks", masks), ("rewards", r), ("values", v),
                         ("log_probs", lp), ("terminations", term)):
                want[k].append(a[e, st])
        data = rb.get_buffer_data()
        assert rb.buffer_size == sum(len(x) for x in want["rewards"]) == len(data["rewards"])
        for k, parts in want.items():
            np.testing.assert_array_equal(data[k], np.concatenate(parts), err_msg=k)
        assert data["observations"].dtype == np.float32 and data["actions"].dtype == np.float32
        assert data["action_masks"].dtype == bool and data["terminations"].dtype == bool
        assert data["observations"].shape[1:] == dims
        with pytest.raises(ValueError, match="generic rows"):
            rb.get_packed()
        rb.reset()
        assert rb.buffer_size == 0 and rb.get_buffer_data()["rewards"].shape == (0,)


def test_rollout_buffer_mixes_packed_and_generic_batches(G):
    """2048 batches are packed as bitboards, a batch with soft observations (or actions that are not one-hot) falls
    back to rows; get_buffer_data concatenates both in store order."""
    ro = G.BatchRunner(9, G.act_randomly).run_actions_batch(8)
    obs, actions, masks, log_probs, values, rewards, terms = ro
    b, t = actions.shape
    onehot = np.eye(4, dtype=np.float32)[actions]
    zeros = np.zeros((b, t), np.float32)
    rb = G.RolloutBuffer(31, (4, 4), 4)
    n1 = rb.store_batch(obs, onehot, masks, rewards, zeros, log_probs, terms)
    assert rb._parts[-1][0] == "packed"
    soft = obs.astype(np.float32) * 0.5
    n2 = rb.store_batch(soft, onehot, masks, rewards, zeros, log_probs, terms)
    assert rb._parts[-1][0] == "generic" and n1 == n2 == rb.buffer_size // 2
    data = rb.get_buffer_data()
    e, st = O.store_batch_indices(terms)
    np.testing.assert_array_equal(data["observations"][:n1], obs[e, st].astype(np.float32))
    np.testing.assert_array_equal(data["observations"][n1:], soft[e, st])
    np.testing.assert_array_equal(data["actions"], np.concatenate([onehot[e, st]] * 2))
    np.testing.assert_array_equal(data["action_masks"], np.concatenate([masks[e, st]] * 2))
    assert data["observations"].shape == (2 * n1, 4, 4, 31)


def test_store_packed_equals_store_batch(G):
    runner = G.BatchRunner(5, G.act_randomly)
    ro = runner.run_packed_batch(48)
    rb1 = G.RolloutBuffer(31, 16, 4)
    kept = rb1.store_packed(ro)
    assert kept == ro.env_steps == rb1.buffer_size
    runner2 = G.BatchRunner(5, G.act_randomly)
    obs, actions, masks, log_probs, values, rewards, terms = runner2.run_actions_batch(48)
    rb2 = G.RolloutBuffer(31, (4, 4), 4)
    rb2.store_batch(obs, np.eye(4, dtype=np.float32)[actions], masks, rewards, np.zeros_like(rewards), log_probs, terms)
    d1, d2 = rb1.get_buffer_data(), rb2.get_buffer_data()
    np.testing.assert_array_equal(d1["observations"].reshape(-1, 496), d2["observations"].reshape(-1, 496))
    assert d2["observations"].shape[1:] == (4, 4, 31)
    for k in ("actions", "action_masks", "rewards", "log_probs", "terminations"):
        np.testing.assert_array_equal(d1[k], d2[k])
    # env-major, first-done-inclusive, as the oracle's restatement of store_batch
    e, s = O.store_batch_indices(terms)
    np.testing.assert_array_equal(d1["rewards"], rewards[e, s])
    assert d1["terminations"].sum() == 48


# ------------------------------------------------------------------------------------- GAE / dataset
def test_ppo_dataset_matches_reference_fixture(G, golden_ppo):
    g = golden_ppo
    n = g["gae_default_rewards"].shape[0]
    buf = {
        "observations": np.zeros((n, 16, 31), np.float32), "actions": np.zeros((n, 4), np.float32),
        "action_masks": np.ones((n, 4), bool), "rewards": g["gae_default_rewards"], "values": g["gae_default_values"],
        "log_probs": np.zeros(n, np.float32), "terminations": g["gae_default_dones"],
    }
    ds = G.PPODataset(buf, gamma=0.99, lambda_gae=0.95)
    adv, ret = ds._compute_gae_returns()
    np.testing.assert_array_equal(adv.numpy(), g["gae_default_adv"])
    np.testing.assert_array_equal(ret.numpy(), g["gae_default_ret"])
    np.testing.assert_allclose(ds.advantages.numpy(), g["gae_default_adv_norm"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(ds.returns.numpy(), g["gae_default_ret_norm"], rtol=1e-5, atol=1e-6)
    assert len(ds) == n and set(ds[0]) == {
        "observations", "actions", "action_masks", "rewards", "values", "log_probs", "terminations", "advantages", "returns"}
    ds2 = G.PPODataset(buf, gamma=0.9, lambda_gae=0.5)
    assert not torch.equal(ds2.advantages, ds.advantages)  # tests/ppo/test_data_loader.py:126-138
    sub = G.PPODataset(buf, max_samples_per_epoch=100, shuffle_on_reset=True)
    assert len(sub) == 100
    first = sub.active_indices.clone()
    sub.reset_epoch()
    assert not torch.equal(first, sub.active_indices)
    loader = G.create_ppo_dataloader(buf, batch_size=64, max_samples_per_epoch=640)
    batch = next(iter(loader))
    assert batch["observations"].shape == (64, 16, 31) and batch["advantages"].shape == (64,)


def test_dataloader_batches_equal_default_collation(G):
    """create_ppo_dataloader fetches a batch with one indexed read per field (PPODataset.__getitems__ + its collate
    function); the batches are those torch's default collation of the per-sample dicts gives, also for a DataLoader
    a caller builds around the dataset with the default collate function."""
    from torch.utils.data import DataLoader

    rng = np.random.default_rng(3)
    n = 1000
    data = {"observations": rng.random((n, 16, 31)).astype(np.float32), "actions": np.eye(4, dtype=np.float32)[rng.integers(0, 4, n)],
            "action_masks": rng.random((n, 4)) < 0.7, "rewards": rng.random(n).astype(np.float32),
            "values": rng.standard_normal(n).astype(np.float32), "log_probs": -rng.random(n).astype(np.float32),
            "terminations": rng.random(n) < 0.02}
    for max_samples in (None, 300):
        torch.manual_seed(1)
        fast = G.create_ppo_dataloader(data, batch_size=64, shuffle=False, drop_last=False, max_samples_per_epoch=max_samples)
        plain = DataLoader(fast.dataset, batch_size=64, shuffle=False, drop_last=False)  # default collate: per-sample path
        assert fast.batch_size == 64 and len(fast) == len(plain) == -(-(max_samples or n) // 64)
        for a, b in zip(fast, plain):
            assert a.keys() == b.keys()
            for k in a:
                assert a[k].dtype == b[k].dtype and torch.equal(a[k], b[k]), k
        item = fast.dataset[5]
        assert item["observations"].shape == (16, 31) and item["advantages"].shape == ()


def test_compute_gae_on_a_real_packed_buffer(G):
    ro = G.BatchRunner(9, G.act_randomly).run_packed_batch(256)
    rb = G.RolloutBuffer(31, 16, 4)
    rb.store_packed(ro)
    p = rb.get_packed()
    values = torch.randn(p["rewards"].shape[0], device="cuda")
    dones = G.ppo.data_loader.meta_to_dones(p["meta"])
    adv, ret = G.compute_gae(p["rewards"], values, dones, normalize=False)
    wa, wr = CO.gae(p["rewards"].cpu().numpy(), values.cpu().numpy(), dones.cpu().numpy(), 0.99, 0.95)
    np.testing.assert_array_equal(adv.cpu().numpy(), wa)
    np.testing.assert_array_equal(ret.cpu().numpy(), wr)
    adv_n, ret_n = G.compute_gae(p["rewards"], values, dones, normalize=True)
    np.testing.assert_allclose(adv_n.cpu().numpy(), O.normalize(wa), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(ret_n.cpu().numpy(), O.normalize(wr), rtol=1e-5, atol=1e-6)


def test_device_minibatches_equal_the_reference_format_dataset(G):
    """DevicePPOBatches (packed buffer, gather kernel) vs PPODataset over get_buffer_data() of the same buffer."""
    ro = G.BatchRunner(21, G.act_randomly).run_packed_batch(96)
    rb = G.RolloutBuffer(31, 16, 4)
    rb.store_packed(ro)
    packed = rb.get_packed()
    packed["values"].copy_(torch.randn_like(packed["values"]))
    ds = G.PPODataset(rb.get_buffer_data(), gamma=0.99, lambda_gae=0.95)
    dl = G.DevicePPOBatches(packed, gamma=0.99, lambda_gae=0.95, batch_size=256, shuffle=True,
                            generator=torch.Generator(device="cuda").manual_seed(3))
    n = rb.buffer_size
    assert len(dl) == n // 256 and dl.total_length == n == len(ds)
    np.testing.assert_allclose(dl.advantages.cpu().numpy(), ds.advantages.numpy(), rtol=1e-5, atol=1e-6)
    seen = []
    for batch in dl:
        assert batch["observations"].shape == (256, 16, 31) and batch["observations"].is_cuda
        seen.append(batch)
    assert len(seen) == len(dl)
    # explicit indices: every field equals the dataset's items
    idx = torch.randperm(n, device="cuda")[:500]
    b = dl.batch(idx)
    items = [ds[int(i)] for i in idx.cpu()]
    for key in ("observations", "action_masks", "log_probs", "values"):
        want = torch.stack([it[key] for it in items])
        assert torch.equal(b[key].cpu(), want), key
    np.testing.assert_array_equal(b["actions"].cpu().numpy(), torch.stack([it["actions"] for it in items]).argmax(-1).numpy())
    np.testing.assert_allclose(b["advantages"].cpu().numpy(), torch.stack([it["advantages"] for it in items]).numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(b["returns"].cpu().numpy(), torch.stack([it["returns"] for it in items]).numpy(), rtol=1e-5, atol=1e-6)
    # bf16 observations for an autocast forward, subset per epoch with reshuffle
    sub = G.DevicePPOBatches(packed, batch_size=64, max_samples_per_epoch=640, shuffle_on_reset=True, obs_dtype=torch.bfloat16)
    assert len(sub) == 10
    first = sub.active_indices.clone()
    sub.reset_epoch()
    assert not torch.equal(first, sub.active_indices)
    batch = next(iter(sub))
    assert batch["observations"].dtype == torch.bfloat16 and float(batch["observations"].float().sum()) == 64 * 16
    # without an explicit generator the subset and the shuffle come from g2048_random_subset, keyed from torch's
    # global generator: distinct positions, every sample of the subset exactly once per epoch, reproducible
    assert len(torch.unique(sub.active_indices)) == 640 and int(sub.active_indices.max()) < n
    torch.manual_seed(77)
    a = G.DevicePPOBatches(packed, batch_size=64, max_samples_per_epoch=640, obs_dtype=None)
    order_a = torch.cat([b["boards"] for b in a])
    torch.manual_seed(77)
    b2 = G.DevicePPOBatches(packed, batch_size=64, max_samples_per_epoch=640, obs_dtype=None)
    assert torch.equal(a.active_indices, b2.active_indices)
    assert torch.equal(order_a, torch.cat([b["boards"] for b in b2]))
    assert torch.equal(torch.sort(order_a).values, torch.sort(packed["boards"][a.active_indices]).values)


# ------------------------------------------------------------------------------------- policy network path
class TinyAgent(torch.nn.Module):
    """Same call signature as the reference's PPOAgent.forward (obs (B,16,31), mask|None)."""

    def __init__(self):
        super().__init__()
        torch.manual_seed(0)
        self.body = torch.nn.Linear(16 * 31, 32)
        self.actor = torch.nn.Linear(32, 4)
        self.critic = torch.nn.Linear(32, 1)

    def forward(self, observations, action_mask=None):
        h = torch.tanh(self.body(observations.reshape(observations.shape[0], -1)))
        logits = self.actor(h)
        if action_mask is not None:
            logits = logits - (1e8 * (1 - action_mask.float()))
        return logits, self.critic(h)


@pytest.mark.parametrize("sample", [True, False])
def test_torch_action_function_in_batch_runner(G, sample):
    agent = TinyAgent()
    fn = G.TorchActionFunction(agent, use_mask=True, sample_actions=sample, device=torch.device("cuda"))
    runner = G.BatchRunner(init_seed=42, act_fn=fn)
    obs, actions, masks, log_probs, values, rewards, terms = runner.run_actions_batch(24)
    b, t = actions.shape
    assert obs.shape == (24, t, 4, 4, 31) and terms[:, -1].all() and np.any(rewards > 0)
    assert np.isfinite(log_probs).all() and np.isfinite(values).all()
    # masked policy never takes an illegal action while the env is alive
    alive = np.concatenate([np.ones((b, 1), bool), ~terms[:, :-1]], axis=1)
    assert masks[np.arange(b)[:, None], np.arange(t)[None, :], actions][alive].all()
    assert not (rewards == -1).any()
    # log-probs / values equal the agent's own evaluate path (tests/ppo/test_log_prob_consistency.py)
    with torch.no_grad():
        flat_obs = torch.from_numpy(obs.reshape(b * t, 16, 31)).float().cuda()
        logits, v = agent(flat_obs, torch.from_numpy(masks.reshape(b * t, 4)).cuda())
        lp = torch.distributions.Categorical(logits=logits).log_prob(torch.from_numpy(actions.reshape(-1)).long().cuda())
    assert np.abs(lp.cpu().numpy().reshape(b, t) - log_probs).max() < 1e-5
    np.testing.assert_allclose(v.cpu().numpy().reshape(b, t), values, rtol=1e-5, atol=1e-6)
    if not sample:
        np.testing.assert_array_equal(actions[alive], logits.argmax(1).cpu().numpy().reshape(b, t)[alive])
    # determinism after re-construction with the same seed (tests/integration/test_ppo_integration.py:249-276)
    again = G.BatchRunner(init_seed=42, act_fn=fn).run_actions_batch(24)
    np.testing.assert_array_equal(again[1], actions)
    # single (un-batched) call keeps the reference's contract
    a, lp1, v1 = fn(G.keys.key(1), obs[0, 0], masks[0, 0])
    assert a.shape == () and masks[0, 0][a.item()]


def test_single_legal_action_is_respected(G):
    agent = TinyAgent()
    for sample in (True, False):
        fn = G.TorchActionFunction(agent, use_mask=True, sample_actions=sample, device=torch.device("cuda"))
        obs = np.zeros((4, 4, 4, 31), np.float32)
        obs[..., 0] = 1
        keys = G.keys.split(G.keys.key(5), 4)
        actions, _, _ = fn(keys, obs, np.eye(4, dtype=bool))
        assert actions.cpu().tolist() == [0, 1, 2, 3]


def test_torch_action_function_autocast_keeps_float32_sampling():
    import g2048

    class Agent(torch.nn.Module):
        def __init__(self):
            super().__init__()
            torch.manual_seed(1)
            self.lin = torch.nn.Linear(496, 5)
            self.seen = None

        def forward(self, obs, mask=None):
            out = self.lin(obs.flatten(1))
            self.seen = out.dtype
            return out[:, :4], out[:, 4:]

    agent = Agent()
    fn = g2048.TorchActionFunction(agent, use_mask=True, device=torch.device("cuda"), autocast_dtype=torch.bfloat16)
    ro = g2048.BatchRunner(init_seed=2, act_fn=fn).run_packed_batch(32)
    assert agent.seen == torch.bfloat16  # the GEMM ran under autocast
    assert ro.log_probs.dtype == torch.float32 and ro.values.dtype == torch.float32
    assert bool(torch.isfinite(ro.log_probs).all()) and ro.env_steps > 0


def test_collect_rollouts_matches_the_trainers_python_loops():
    """ppo_trainer.py:185-227 on the reference-format arrays vs the packed path."""
    import g2048
    from g2048.ppo import collect_rollouts

    ref_runner = g2048.BatchRunner(init_seed=6, act_fn=g2048.act_randomly)
    ref_buffer = g2048.RolloutBuffer(31, 16, 4)
    want_r, want_l = [], []
    for _ in range(3):
        obs, actions, masks, log_probs, values, rewards, terms = ref_runner.run_actions_batch(40)
        b, t = obs.shape[:2]
        onehot = np.zeros((b, t, 4), np.float32)
        onehot[np.arange(b)[:, None], np.arange(t)[None, :], actions] = 1.0
        ref_buffer.store_batch(observations=obs, actions=onehot, action_masks=masks, rewards=rewards, values=values,
                               log_probs=log_probs, terminations=terms)
        for e in range(b):
            want_r.append(np.max(rewards[e]))
            want_l.append(np.argmax(terms[e]) + 1 if np.any(terms[e]) else t)
    for compact in (False, True):
        runner = g2048.BatchRunner(init_seed=6, act_fn=g2048.act_randomly, compact_live=compact)
        buffer = g2048.RolloutBuffer(31, 16, 4)
        out = collect_rollouts(runner, buffer, 40, 3)
        np.testing.assert_array_equal(out["episode_rewards"], np.asarray(want_r, np.float32))
        np.testing.assert_array_equal(out["episode_lengths"], np.asarray(want_l))
        assert out["total_episodes"] == 120 and out["timesteps"] == ref_buffer.buffer_size == buffer.buffer_size
        a, b2 = ref_buffer.get_buffer_data(), buffer.get_buffer_data()
        for k in a:
            np.testing.assert_array_equal(a[k], b2[k])


def test_pinned_outputs_are_the_same_arrays_in_page_locked_memory(G):
    """BatchRunner(pinned_outputs=True): run_actions_batch returns the same values; the large arrays are page-locked
    memory (one transfer each), ordinary numpy arrays to the caller -- writable, sliceable, alive after the runner."""
    plain = G.BatchRunner(init_seed=4, act_fn=G.act_randomly).run_actions_batch(300)
    runner = G.BatchRunner(init_seed=4, act_fn=G.act_randomly, pinned_outputs=True)
    pinned = runner.run_actions_batch(300)
    del runner
    for a, b in zip(plain, pinned):
        assert (a is None and b is None) or (a.dtype == b.dtype and a.shape == b.shape and np.array_equal(a, b))
    obs = pinned[0]
    assert obs.nbytes >= 1 << 20 and torch.from_numpy(obs).is_pinned() and not torch.from_numpy(plain[0]).is_pinned()
    obs[0, 0] = False  # writable
    assert not obs[0, 0].any() and plain[0][0, 0].any()


def test_key_chain_generated_ahead_is_the_same_chain(G):
    """KeyChain produces its sub keys on a side stream, a block ahead of the consumer: any sequence of peeks and
    consumes sees the keys of one synchronous generation, also when chains are created and dropped in a loop while
    their last generation is still in flight (the dropped chain's memory must not be reused under it)."""
    from g2048 import engine as E
    from g2048.keys import KeyChain

    rng = np.random.default_rng(0)
    for mode in (0, 1):
        ref = E.chain_advance(E.words_tensor([0, 7], "cuda"), mode, 150_000).cpu()
        chain, pos = KeyChain(7, mode), 0
        for _ in range(40):
            n = int(rng.integers(1, 9000))
            assert torch.equal(chain.peek(n).cpu(), ref[pos:pos + n])
            k = int(rng.integers(0, n + 1))
            chain.consume(k)
            pos += k
        assert chain.position == pos
        want_key = E.words_tensor([0, 7], "cuda")
        E.chain_advance(want_key, mode, pos)
        assert np.array_equal(chain.key, E.words_numpy(want_key))
    # churn: a fresh chain per iteration, dropped with a generation in flight; every one must still be right
    first = E.chain_advance(E.words_tensor([0, 9], "cuda"), 1, 300).cpu()
    for _ in range(200):
        got = KeyChain(9, 1).peek(300)
        junk = torch.full((64,), -1, dtype=torch.int32, device="cuda")  # takes freed blocks of the main stream's pool
        assert torch.equal(got.cpu(), first)
        del junk
