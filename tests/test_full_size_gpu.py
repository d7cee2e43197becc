"""BASELINE.json's full sizes on the GPU, checked through properties that do not need a full-size oracle run
(conservation between per-env outputs and on-device reductions, terminal final boards, idempotence, shard
invariance, episode independence of GAE, round trips) plus bit-exact oracle parity on samples cut out of the
full-size batch -- the oracle can play any env range of a batch because keys derive from global env indices."""
import numpy as np
import pytest
import torch

from oracle import c_oracle as CO
from oracle import pgx2048_oracle as O

pytestmark = pytest.mark.gpu

MODE = 1  # partitionable (jax 0.5.3 default)


@pytest.fixture(scope="module")
def E():
    from g2048 import engine

    assert torch.cuda.is_available()
    return engine


def _play(E, policy, seed, n, lo=0, cnt=None, max_steps=2048):
    cnt = n if cnt is None else cnt
    key = E.words_tensor(list(E.key_words(seed)), "cuda")
    subs = E.chain_advance(key, MODE, 1 + 2 * max_steps)
    return E.play(policy, subs, n, lo, cnt, MODE, per_env=True)


@pytest.mark.parametrize("policy,n,seed", [(0, 1 << 21, 2048), (1, 1 << 20, 0)])  # C5 per-GPU shard; C2
def test_play_to_termination_at_full_size(E, policy, n, seed):
    out = _play(E, policy, seed, n)
    stats = E.play_stats_dict(out["stats"])
    lengths = out["lengths"].cpu().numpy().astype(np.int64)
    scores = out["scores"].cpu().numpy().astype(np.int64)
    boards = E.boards_numpy(out["final_boards"])
    # conservation: the on-device reduction agrees with the per-env outputs
    assert stats["episodes"] == n and stats["cut_short"] == 0 and stats["overflowed"] == 0
    assert stats["env_steps"] == int(lengths.sum()) and stats["longest"] == int(lengths.max())
    assert stats["score_sum"] == int(scores.sum())
    hist = np.bincount(boards.max(axis=1), minlength=32)
    assert {int(k): int(v) for k, v in stats["max_tile_hist"].items()} == {1 << e: int(c) for e, c in enumerate(hist) if c}
    # every episode ran to a terminal board: no action is legal on it; rewards are multiples of 4
    assert not O.exact_legal(boards).any()
    assert (scores % 4 == 0).all() and lengths.min() >= 1
    # idempotence: the same seed gives the same batch
    again = _play(E, policy, seed, n)
    for k in ("final_boards", "lengths", "scores"):
        assert torch.equal(out[k], again[k])
    # shard invariance: two halves with global env indices are the halves of the whole
    half = n // 2
    for lo in (0, half):
        part = _play(E, policy, seed, n, lo, half)
        for k in ("final_boards", "lengths", "scores"):
            assert torch.equal(part[k], out[k][lo:lo + half])
    # bit-exact oracle parity on 4 096 envs cut out of the middle of the batch
    lo = n // 2 + 12345
    want = CO.play(seed, n, policy, MODE, env_lo=lo, env_hi=lo + 4096, max_steps=2048)
    np.testing.assert_array_equal(boards[lo:lo + 4096], want["final_boards"])
    np.testing.assert_array_equal(lengths[lo:lo + 4096], want["lengths"])
    np.testing.assert_array_equal(scores[lo:lo + 4096], want["scores"])


def test_flat_gae_at_c4_size(E):
    """C4: ~7.7e7 steps (256 K envs x ~300 steps).  Episodes are independent, so (1) a window that starts after a
    done and ends on one equals the oracle on that window alone, (2) swapping two blocks of whole episodes swaps the
    outputs, (3) a done step has adv = r - V exactly; and the moment block equals double-precision sums."""
    n = 77_000_000
    gen = torch.Generator(device="cuda")
    gen.manual_seed(7)
    r = (torch.randint(0, 64, (n,), device="cuda", generator=gen) * 4).float()
    v = torch.randn(n, device="cuda", generator=gen)
    d = (torch.rand(n, device="cuda", generator=gen) < 1 / 300).to(torch.uint8)
    d[-1] = 1
    adv, ret, mom = E.gae_flat(r, v, d, 0.99, 0.95)
    ends = torch.nonzero(d).flatten()
    # (1) windows of whole episodes against the oracle, at several places incl. across many tile borders
    for k0 in (0, 1000, len(ends) // 2, len(ends) - 400):
        a, b = (0 if k0 == 0 else int(ends[k0 - 1]) + 1), int(ends[k0 + 300]) + 1
        wa, wr = CO.gae(r[a:b].cpu().numpy(), v[a:b].cpu().numpy(), d[a:b].cpu().numpy(), 0.99, 0.95)
        np.testing.assert_array_equal(adv[a:b].cpu().numpy(), wa)
        np.testing.assert_array_equal(ret[a:b].cpu().numpy(), wr)
    # (3) at a done step the recurrence restarts
    assert torch.equal(adv[ends], r[ends] - v[ends])
    assert torch.equal(ret, adv + v)
    # moments
    m = mom.cpu().numpy()
    assert m[0] == n
    np.testing.assert_allclose(m[1], float(adv.double().sum()), rtol=1e-9, atol=1e-3)
    np.testing.assert_allclose(m[2], float((adv.double() ** 2).sum()), rtol=1e-9)
    np.testing.assert_allclose(m[4], float((ret.double() ** 2).sum()), rtol=1e-9)
    # (2) swap the two halves at an episode border
    cut = int(ends[len(ends) // 2]) + 1
    swap = lambda x: torch.cat([x[cut:], x[:cut]])  # noqa: E731
    adv2, ret2, _ = E.gae_flat(swap(r), swap(v), swap(d), 0.99, 0.95)
    assert torch.equal(adv2, swap(adv)) and torch.equal(ret2, swap(ret))
    del adv2, ret2
    # the three kernel generations agree bit for bit at this size too
    for entry in ("g2048_gae_flat_tiled", "g2048_gae_flat_pipelined"):
        a3, r3, _ = E.gae_flat(r, v, d, 0.99, 0.95, entry=entry)
        assert torch.equal(a3, adv) and torch.equal(r3, ret), entry


def test_time_major_gae_and_observations_at_c3_size(E):
    """C3: 128 steps x 65 536 envs.  Time-major GAE of env e equals the flat GAE of its column; the one-hot
    expansion round-trips through pack_obs and has exactly one 1 per cell."""
    t_steps, b = 128, 1 << 16
    gen = torch.Generator(device="cuda")
    gen.manual_seed(3)
    rr = (torch.randint(0, 32, (t_steps, b), device="cuda", generator=gen) * 4).float()
    vv = torch.randn(t_steps, b, device="cuda", generator=gen)
    dd = torch.rand(t_steps, b, device="cuda", generator=gen) < 1 / 100
    meta = (dd.to(torch.uint8) << 6) | torch.randint(0, 64, (t_steps, b), device="cuda", generator=gen).to(torch.uint8)
    adv, ret, mom = E.gae_time_major(rr, vv, meta, t_steps, b, None, 0.99, 0.95)
    # env-major flattening with a done forced at every env's last step = independent columns
    d_flat = dd.t().contiguous().to(torch.uint8)
    d_flat[:, -1] = 1
    fa, fr, _ = E.gae_flat(rr.t().contiguous().view(-1), vv.t().contiguous().view(-1), d_flat.view(-1), 0.99, 0.95)
    assert torch.equal(adv.t().contiguous().view(-1), fa) and torch.equal(ret.t().contiguous().view(-1), fr)
    assert mom[0].item() == t_steps * b
    boards = torch.randint(0, 1 << 62, (1 << 20,), dtype=torch.int64, device="cuda", generator=gen)
    obs = E.expand_obs(boards, torch.float32)
    assert torch.equal(E.pack_obs(obs), boards)
    assert float(obs.sum()) == boards.numel() * 16 and bool((obs.sum(dim=2) == 1).all())


def test_recorded_rollout_and_buffer_at_scale():
    """262 144 envs recorded to termination (C4's rollout): the compacted buffer keeps exactly the live steps, every
    episode ends with its only done, and the kept boards are the recorded pre-step boards in env-major order."""
    import g2048
    from g2048 import engine as E

    n = 1 << 18
    runner = g2048.BatchRunner(init_seed=4, act_fn=g2048.act_drul)
    ro = runner.run_packed_batch(n)
    lengths = ro.lengths().long()
    assert int(lengths.sum()) == ro.env_steps and int(lengths.max()) == ro.t_steps
    buf = g2048.RolloutBuffer(31, 16, 4)
    kept = buf.store_packed(ro)
    packed = buf.get_packed()
    assert kept == ro.env_steps == packed["boards"].shape[0]
    dones = E.meta_dones(packed["meta"]).bool()
    ends = torch.cumsum(lengths, 0) - 1
    assert int(dones.sum()) == n and bool(dones[ends].all())
    # env e's episode occupies [ends[e] - len + 1, ends[e]]; compare three envs' boards with the time-major record
    for e in (0, n // 2, n - 1):
        length = int(lengths[e])
        start = int(ends[e]) - length + 1
        assert torch.equal(packed["boards"][start:start + length], ro.boards[:length, e])
        assert torch.equal(packed["rewards"][start:start + length], ro.rewards[:length, e])
