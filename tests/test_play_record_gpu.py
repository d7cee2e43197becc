"""The recording form of the persistent table kernel (g2048_play_record + g2048_play_record_compact): the env-major
flat buffer it produces must equal, bit for bit, what the lock-step recorder + RolloutBuffer.store_packed produce
(src/runs/batch_runner.py:117-154 + src/ppo/rollout_buffer.py:164-187), and the numpy oracle's trajectories."""
import numpy as np
import pytest
import torch

from oracle import pgx2048_oracle as O

pytestmark = pytest.mark.gpu

MODES = [0, 1]


@pytest.fixture(scope="module")
def G():
    import g2048

    return g2048


@pytest.fixture(scope="module")
def E():
    from g2048 import engine

    return engine


def _policy_fn(G, policy):
    return G.act_randomly if policy == 0 else G.act_drul


def _assert_flat_equals_packed(G, flat, seed, policy, mode, n, shard=None):
    runner = G.BatchRunner(init_seed=seed, act_fn=_policy_fn(G, policy), rng_mode=mode, shard=shard)
    ro = runner.run_packed_batch(n)
    buf = G.RolloutBuffer(31, 16, 4)
    buf.store_packed(ro)
    want = buf.get_packed()
    assert flat.env_steps == ro.env_steps == want["boards"].shape[0]
    for name in ("boards", "meta", "rewards", "log_probs", "values"):
        assert torch.equal(getattr(flat, name), want[name]), name
    lengths = ro.lengths()
    assert torch.equal(flat.lengths, lengths)
    assert torch.equal(flat.offsets[:-1], torch.cumsum(lengths.long(), 0) - lengths.long())
    assert int(flat.offsets[-1]) == flat.env_steps
    assert torch.equal(flat.final_boards, ro.final_boards)
    assert torch.equal(flat.max_rewards, ro.rewards[: ro.t_steps].max(dim=0).values)
    assert flat.t_steps == ro.t_steps
    return runner


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("policy", [0, 1])
@pytest.mark.parametrize("n", [1, 37, 1000, 40001])
def test_flat_rollout_equals_lock_step_recorder_plus_store_packed(G, mode, policy, n):
    seed = 11 + n
    runner = G.BatchRunner(init_seed=seed, act_fn=_policy_fn(G, policy), rng_mode=mode)
    flat = runner.run_flat_batch(n)
    other = _assert_flat_equals_packed(G, flat, seed, policy, mode, n)
    np.testing.assert_array_equal(runner.key, other.key)  # both consumed 1 + 2 T sub keys
    # scores = the sum of every env's rewards; every episode ends with its only done
    seg = torch.repeat_interleave(torch.arange(n, device="cuda"), flat.lengths.long())
    sums = torch.zeros(n, dtype=torch.float64, device="cuda").index_add_(0, seg, flat.rewards.double())
    assert torch.equal(sums.long(), flat.scores.long())
    done = (flat.meta >> 6) & 1
    ends = flat.offsets[1:] - 1
    assert int(done.sum()) == n and bool(done[ends].all())


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("policy,name", [(0, "random"), (1, "drul")])
def test_flat_rollout_matches_numpy_oracle(G, E, mode, policy, name):
    n, seed = 48, 77
    ref = O.rollout(O.KeyChain(seed, mode), n, name)
    flat = G.BatchRunner(init_seed=seed, act_fn=_policy_fn(G, policy), rng_mode=mode).run_flat_batch(n)
    length, score, _ = O.episode_summary(ref)
    np.testing.assert_array_equal(flat.lengths.cpu().numpy(), length)
    np.testing.assert_array_equal(flat.scores.cpu().numpy(), score)
    boards = E.boards_numpy(flat.boards)
    meta = flat.meta.cpu().numpy()
    rewards = flat.rewards.cpu().numpy()
    offs = flat.offsets.cpu().numpy()
    ref_boards = np.stack(ref["boards"])            # (T, n, 16) pre-step
    ref_actions = np.stack(ref["actions"])
    ref_masks = np.stack(ref["masks"])
    ref_rewards = np.stack([s.rewards[:, 0] for s in ref["states"]])
    for e in range(n):
        sl = slice(offs[e], offs[e + 1])
        k = int(length[e])
        np.testing.assert_array_equal(boards[sl], ref_boards[:k, e])
        np.testing.assert_array_equal(meta[sl] & 3, ref_actions[:k, e])
        np.testing.assert_array_equal(((meta[sl, None] >> (2 + np.arange(4))) & 1).astype(bool), ref_masks[:k, e])
        np.testing.assert_array_equal(rewards[sl], ref_rewards[:k, e])
    np.testing.assert_array_equal(E.boards_numpy(flat.final_boards), ref["states"][-1].board)


def test_flat_rollout_is_shard_invariant(G):
    n, seed = 5000, 3
    full = G.BatchRunner(init_seed=seed, act_fn=G.act_randomly).run_flat_batch(n)
    parts = [G.BatchRunner(init_seed=seed, act_fn=G.act_randomly, shard=(r, 3)).run_flat_batch(n) for r in range(3)]
    for name in ("boards", "meta", "rewards", "log_probs", "lengths", "final_boards", "scores"):
        assert torch.equal(torch.cat([getattr(p, name) for p in parts]), getattr(full, name)), name


def test_arena_too_small_is_reported_and_retried(G, E):
    """A lane takes a new env only while a whole episode still fits into its arena region; with a mean-length hint of
    1 step every lane retires after its first episode, the kernel reports episodes < n, and the runner retries with a
    larger arena until the batch fits."""
    n = 200_000
    key = E.words_tensor([0, 9], "cuda")
    subs = E.chain_advance(key, 1, 1 + 2 * 2048)
    rec = E.play_record(0, subs, n, 0, n, 1, mean_steps=1)
    st = E.play_stats_dict(rec["stats"])
    assert 0 < st["episodes"] < n and st["cut_short"] == 0
    played = rec["lengths"] > 0
    assert int(played.sum()) == st["episodes"] and int(rec["lengths"].sum()) == st["env_steps"]
    runner = G.BatchRunner(init_seed=9, act_fn=G.act_randomly)
    runner._mean_steps[0] = 1
    flat = runner.run_flat_batch(n)
    assert flat.summary["episodes"] == n and int((flat.lengths > 0).sum()) == n
    want = E.play(0, subs, n, 0, n, 1)
    assert torch.equal(flat.lengths, want["lengths"]) and torch.equal(flat.final_boards, want["final_boards"])


def test_compaction_into_estimated_room(G, E):
    """run_flat_batch launches the compaction before it knows the batch's total: into arrays sized from an estimate.
    Too little room: the steps that fit are written, nothing beyond, the per-env maxima are complete, and the runner
    compacts a second time; more than enough: the rollout's arrays are prefixes of the roomier ones.  Either way the
    buffer is the one an exact-size compaction produces."""
    n = 6000
    subs = E.chain_advance(E.words_tensor([0, 31], "cuda"), 1, 1 + 2 * 2048)
    rec = E.play_record(1, subs, n, 0, n, 1)
    offsets = E.exclusive_scan(rec["lengths"])
    total = int(offsets[-1])
    exact = E.play_record_compact(rec, offsets, total)
    for room in (0, 1, total // 3, total - 1, total + 5000):
        # guard band: the arrays are really `room + 64` long and pre-filled; only [0, min(room, total)) may change
        got = E.play_record_compact(rec, offsets, room)
        keep = min(room, total)
        for k in ("boards", "meta", "rewards", "log_probs", "values"):
            assert got[k].shape[0] == room and torch.equal(got[k][:keep], exact[k][:keep]), (k, room)
        assert torch.equal(got["max_rewards"], exact["max_rewards"])
    # through the raw entry point with guard bands behind a short capacity
    from g2048 import _native as N

    room = total // 2
    band = {k: torch.full((room + 64,), 7, dtype=exact[k].dtype, device="cuda") for k in ("boards", "meta", "rewards", "log_probs", "values")}
    N.call("g2048_play_record_compact", 1, N.ptr(rec["arena_boards"]), N.ptr(rec["arena_meta"]), N.ptr(rec["env_slot"]),
           N.ptr(rec["lengths"]), N.ptr(offsets), n, 0, room, N.ptr(band["boards"]), N.ptr(band["meta"]), N.ptr(band["rewards"]),
           N.ptr(band["log_probs"]), N.ptr(band["values"]), None, N.stream_ptr())
    for k, v in band.items():
        assert torch.equal(v[:room], exact[k][:room]) and bool((v[room:] == 7).all()), k

    # the runner: a far too small estimate (second compaction) and a generous one give the same rollout
    want = G.BatchRunner(init_seed=31, act_fn=G.act_drul).run_flat_batch(n)
    for guess in (16, 4000):
        runner = G.BatchRunner(init_seed=31, act_fn=G.act_drul)
        runner._mean_steps[1] = guess
        got = runner.run_flat_batch(n)
        assert got.env_steps == want.env_steps == total
        for k in ("boards", "meta", "rewards", "log_probs", "values", "lengths", "offsets", "final_boards", "scores", "max_rewards"):
            assert torch.equal(getattr(got, k), getattr(want, k)), (k, guess)
        assert got.boards.is_contiguous() and got.boards.shape[0] == total


def test_store_flat_feeds_the_training_pipeline(G):
    """collect_rollouts on a built-in policy goes through run_flat_batch / store_flat and yields the same buffer and
    episode statistics as the packed path (see test_collect_rollouts_matches_the_trainers_python_loops)."""
    from g2048.ppo import collect_rollouts

    ra = G.BatchRunner(init_seed=6, act_fn=G.act_drul)
    ba = G.RolloutBuffer(31, 16, 4)
    out = collect_rollouts(ra, ba, 300, 2)
    rb = G.BatchRunner(init_seed=6, act_fn=G.act_drul)
    bb = G.RolloutBuffer(31, 16, 4)
    lens, rews = [], []
    for _ in range(2):
        ro = rb.run_packed_batch(300)
        bb.store_packed(ro)
        lens.append(ro.lengths().cpu().numpy())
        rews.append(ro.rewards.max(dim=0).values.cpu().numpy())
    np.testing.assert_array_equal(out["episode_lengths"], np.concatenate(lens))
    np.testing.assert_array_equal(out["episode_rewards"], np.concatenate(rews))
    pa, pb = ba.get_packed(), bb.get_packed()
    for k in pa:
        assert torch.equal(pa[k], pb[k]), k
    da, db = ba.get_buffer_data(), bb.get_buffer_data()
    for k in da:
        np.testing.assert_array_equal(da[k], db[k])
    n_batches = sum(1 for _ in G.DevicePPOBatches(pa, batch_size=512))
    assert n_batches == ba.buffer_size // 512  # drop_last
