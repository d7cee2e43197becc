"""CUDA-graph form of the network-policy rollout loop: one captured step
(observation -> forward -> sample + env.step + record) replayed, against the eager per-step loop."""
import pytest
import torch

pytestmark = pytest.mark.gpu


class TinyAgent(torch.nn.Module):
    """forward(obs (B,16,31), mask|None) -> (logits (B,4), values (B,1)), the reference agent's call shape."""

    def __init__(self):
        super().__init__()
        torch.manual_seed(5)
        self.body = torch.nn.Sequential(torch.nn.Flatten(), torch.nn.Linear(496, 64), torch.nn.Tanh())
        self.actor = torch.nn.Linear(64, 4)
        self.critic = torch.nn.Linear(64, 1)

    def forward(self, obs, mask=None):
        h = self.body(obs)
        return self.actor(h), self.critic(h)


def _runner(graph: bool, use_mask=True, seed=9):
    import g2048
    fn = g2048.TorchActionFunction(TinyAgent(), use_mask=use_mask, sample_actions=True, device=torch.device("cuda"))
    return g2048.BatchRunner(init_seed=seed, act_fn=fn, cuda_graph=graph)


@pytest.mark.parametrize("batch", [7, 300])
def test_graph_replay_gives_the_eager_records(batch):
    eager, graphed = _runner(False), _runner(True)
    for _ in range(2):  # the second batch continues the key chain and reuses the captured graph
        a, b = eager.run_packed_batch(batch), graphed.run_packed_batch(batch)
        assert a.t_steps == b.t_steps and a.env_steps == b.env_steps
        for name in ("boards", "meta", "rewards", "log_probs", "values", "final_boards", "final_status"):
            assert torch.equal(getattr(a, name), getattr(b, name)), name
        assert (eager.key == graphed.key).all()
    assert len(graphed._graphs) == 1


def test_a_new_action_function_gets_a_new_graph():
    import g2048
    graphed = _runner(True)
    a = graphed.run_packed_batch(16)
    other = TinyAgent()
    with torch.no_grad():
        other.actor.weight.mul_(-3.0)
    graphed.act_fn = g2048.TorchActionFunction(other, use_mask=True, sample_actions=True, device=torch.device("cuda"))
    eager = g2048.BatchRunner(init_seed=9, act_fn=graphed.act_fn)
    eager.load_state_dict(graphed.state_dict())  # same chain position as the graphed runner
    b, c = graphed.run_packed_batch(16), eager.run_packed_batch(16)
    assert torch.equal(b.meta, c.meta) and torch.equal(b.boards, c.boards)
    assert a.t_steps > 0


def test_graph_path_serves_the_reference_api():
    eager, graphed = _runner(False, use_mask=False), _runner(True, use_mask=False)
    out_a, out_b = eager.run_actions_batch(16), graphed.run_actions_batch(16)
    assert len(out_a) == len(out_b) == 7
    for x, y in zip(out_a, out_b):
        assert x.shape == y.shape and (x == y).all()
    assert out_b[6][:, -1].all()  # terminations[:, -1]


def test_policy_step_at_matches_policy_step():
    from g2048 import engine as E
    mode, n, steps = E.RNG_PARTITIONABLE, 500, 5
    subs = E.chain_advance(E.words_tensor([0, 1], "cuda"), mode, 1 + 2 * steps)
    b0, s0 = E.env_init(subs[0], n, 0, n, mode)
    b1, s1 = b0.clone(), s0.clone()
    torch.manual_seed(0)
    logits = torch.randn(steps, n, 4, device="cuda")
    values = torch.randn(steps, n, device="cuda")
    rec = lambda dt: torch.zeros((steps, n), dtype=dt, device="cuda")  # noqa: E731
    ra = [rec(torch.int64), rec(torch.uint8), rec(torch.float32), rec(torch.float32), rec(torch.float32)]
    rb = [rec(torch.int64), rec(torch.uint8), rec(torch.float32), rec(torch.float32), rec(torch.float32)]
    step_index = torch.zeros((), dtype=torch.int32, device="cuda")
    chunk = subs[1:].contiguous()
    for t in range(steps):
        E.policy_step(b0, s0, logits[t], values[t], True, True, False, subs[1 + 2 * t], subs[2 + 2 * t], n, 0, mode,
                      *(x[t] for x in ra))
        E.policy_step_at(b1, s1, logits[t], values[t], True, True, False, chunk, step_index, n, 0, mode, *rb)
        E.counter_add(step_index, 1)
    assert int(step_index.item()) == steps
    assert torch.equal(b0, b1) and torch.equal(s0, s1)
    for x, y in zip(ra, rb):
        assert torch.equal(x, y)


def test_graph_needs_the_network_on_the_gpu():
    import g2048
    fn = g2048.TorchActionFunction(TinyAgent(), device=torch.device("cpu"))
    runner = g2048.BatchRunner(init_seed=1, act_fn=fn, cuda_graph=True)
    with pytest.raises(ValueError, match="network on the runner's GPU"):
        runner.run_packed_batch(4)


# ---------------------------------------------------------------------------- fixed-horizon mode (config C3)
def _fixed(graph: bool, n=300, seed=4):
    import g2048
    fn = g2048.TorchActionFunction(TinyAgent(), use_mask=True, sample_actions=True, device=torch.device("cuda"))
    return g2048.FixedHorizonRunner(init_seed=seed, act_fn=fn, batch_size=n, cuda_graph=graph)


def test_fixed_horizon_rollout_is_the_engine_loop_with_auto_reset():
    """collect(T) == env_init + T x (expand_obs, forward, policy_step with auto_reset) on the reference's key chain."""
    from g2048 import engine as E
    n, t_steps = 300, 40
    runner = _fixed(False, n)
    fn = runner.act_fn
    ro = runner.collect(t_steps)
    mode = E.RNG_PARTITIONABLE
    subs = E.chain_advance(E.words_tensor(list(E.key_words(4)), "cuda"), mode, 1 + 2 * t_steps)
    boards, status = E.env_init(subs[0], n, 0, n, mode)
    for t in range(t_steps):
        assert torch.equal(ro.boards[t], boards)
        logits, values = fn.forward_logits(E.expand_obs(boards, torch.float32))
        rm = torch.empty(n, dtype=torch.uint8, device="cuda")
        rr = torch.empty(n, dtype=torch.float32, device="cuda")
        E.policy_step(boards, status, logits, values, True, True, True, subs[1 + 2 * t], subs[2 + 2 * t], n, 0, mode, None, rm, rr)
        assert torch.equal(ro.meta[t], rm) and torch.equal(ro.rewards[t], rr) and torch.equal(ro.values[t], values)
    assert torch.equal(ro.final_boards, boards) and torch.equal(ro.final_status, status)
    dones = (ro.meta >> 6) & 1
    assert ro.boards.shape == (t_steps, n) and int(dones.sum()) >= 0


def test_fixed_horizon_graph_equals_eager_and_envs_persist():
    eager, graphed = _fixed(False), _fixed(True)
    total_done = 0
    for _ in range(3):  # state and key chain carry over from one collect() to the next
        a, b = eager.collect(64), graphed.collect(64)
        for name in ("boards", "meta", "rewards", "log_probs", "values", "final_boards", "final_status"):
            assert torch.equal(getattr(a, name), getattr(b, name)), name
        total_done += int(((a.meta >> 6) & 1).sum())
        assert torch.equal(a.final_boards, eager.boards)
    assert total_done > 0, "192 steps of 300 envs must finish (and auto-reset) some episodes"
    assert (eager.key == graphed.key).all()


def test_fixed_horizon_gae_bootstraps_running_episodes():
    import numpy as np
    from oracle import pgx2048_oracle as O
    runner = _fixed(False, n=64)
    ro = runner.collect(16)
    boot = runner.bootstrap_values()
    adv, ret, _ = ro.gae(0.99, 0.95, bootstrap=boot)
    r, v = ro.rewards.cpu().numpy(), ro.values.cpu().numpy()
    d = ((ro.meta >> 6) & 1).cpu().numpy().astype(bool)
    b = boot.cpu().numpy()
    want = np.zeros_like(r)
    g32, gl32 = np.float32(0.99), np.float32(0.99 * 0.95)
    for e in range(64):  # the reference recurrence (data_loader.py:103-130) per env, seeded with the bootstrap value
        last_v, last_g = b[e], np.float32(0)
        for t in range(15, -1, -1):
            if d[t, e]:
                last_v, last_g = np.float32(0), np.float32(0)
            delta = np.float32(np.float32(r[t, e] + np.float32(g32 * last_v)) - v[t, e])
            last_g = np.float32(delta + np.float32(gl32 * last_g))
            want[t, e] = last_g
            last_v = v[t, e]
    np.testing.assert_array_equal(adv.cpu().numpy(), want)
    np.testing.assert_array_equal(ret.cpu().numpy(), want + v)
    assert O is not None


def test_fixed_horizon_checkpoint_resumes_env_state_and_key_chain():
    a = _fixed(False, n=100)
    a.collect(20)
    saved = a.state_dict()
    want = a.collect(20)
    b = _fixed(True, n=100, seed=99)  # different seed, then overwritten by the checkpoint
    b.collect(7)
    b.load_state_dict(saved)
    got = b.collect(20)
    for name in ("boards", "meta", "rewards", "log_probs", "values", "final_boards"):
        assert torch.equal(getattr(want, name), getattr(got, name)), name
    with pytest.raises(ValueError):
        _fixed(False, n=50).load_state_dict(saved)


# ---------------------------------------------------------------------------- live-env compaction
class RowwiseAgent(torch.nn.Module):
    """Every output row is computed from its input row with a fixed operation order (elementwise product and a
    row sum, no GEMM), so a row's logits do not depend on which other rows are in the batch."""

    def __init__(self):
        super().__init__()
        g = torch.Generator().manual_seed(11)
        self.w = torch.nn.Parameter(torch.randn(5, 496, generator=g) * 0.3)

    def forward(self, obs, mask=None):
        x = obs.reshape(obs.shape[0], 1, 496)
        out = (x * self.w.unsqueeze(0)).sum(dim=2)
        return out[:, :4], out[:, 4:5]


def test_feeding_only_live_envs_keeps_their_trajectories():
    import g2048
    make = lambda compact: g2048.BatchRunner(  # noqa: E731
        init_seed=21, act_fn=g2048.TorchActionFunction(RowwiseAgent(), use_mask=True, device=torch.device("cuda")),
        compact_live=compact)
    full, compact = make(False), make(True)
    for _ in range(2):
        a, b = full.run_packed_batch(200), compact.run_packed_batch(200)
        assert a.t_steps == b.t_steps and a.env_steps == b.env_steps
        assert torch.equal(a.final_boards, b.final_boards) and torch.equal(a.final_status, b.final_status)
        la = a.lengths().long()
        assert torch.equal(la, b.lengths().long())
        alive = torch.arange(a.t_steps, device="cuda").unsqueeze(1) < la.unsqueeze(0)  # (T, n): steps up to the env's end
        for name in ("boards", "meta", "rewards", "log_probs", "values"):
            x, y = getattr(a, name), getattr(b, name)
            assert torch.equal(x[alive], y[alive]), name
            assert bool((y[~alive] == 0).all()), name  # untouched slots
        ba, bb = g2048.RolloutBuffer(31, 16, 4), g2048.RolloutBuffer(31, 16, 4)
        assert ba.store_packed(a) == bb.store_packed(b)
        pa, pb = ba.get_packed(), bb.get_packed()
        for k in pa:
            assert torch.equal(pa[k], pb[k]), k
        assert (full.key == compact.key).all()
    # the reference-format call is unaffected by the flag
    out_a, out_b = full.run_actions_batch(8), compact.run_actions_batch(8)
    for x, y in zip(out_a, out_b):
        assert (x == y).all()


@pytest.mark.parametrize("policy", ["random", "drul"])
def test_recorded_rollout_of_live_envs_only(policy):
    import g2048
    act = g2048.act_randomly if policy == "random" else g2048.act_drul
    full = g2048.BatchRunner(init_seed=33, act_fn=act)
    live = g2048.BatchRunner(init_seed=33, act_fn=act, compact_live=True)
    for batch in (77, 3000):
        a, b = full.run_packed_batch(batch), live.run_packed_batch(batch)
        assert a.t_steps == b.t_steps and a.env_steps == b.env_steps
        assert torch.equal(a.final_boards, b.final_boards) and torch.equal(a.final_status, b.final_status)
        la = a.lengths().long()
        assert torch.equal(la, b.lengths().long())
        alive = torch.arange(a.t_steps, device="cuda").unsqueeze(1) < la.unsqueeze(0)
        for name in ("boards", "meta", "rewards") + (("log_probs",) if policy == "random" else ()):
            x, y = getattr(a, name), getattr(b, name)
            assert torch.equal(x[alive], y[alive]), name
            assert bool((y[~alive] == 0).all()), name
        assert (full.key == live.key).all()


def test_eval_batch_keeps_only_results_and_matches_the_recorded_run():
    import g2048
    make = lambda: g2048.BatchRunner(  # noqa: E731
        init_seed=8, act_fn=g2048.TorchActionFunction(RowwiseAgent(), use_mask=True, sample_actions=False, device=torch.device("cuda")))
    rec, ev = make(), make()
    for batch in (33, 400):
        ro = rec.run_packed_batch(batch)
        out = ev.run_eval_batch(batch)
        la = ro.lengths()
        assert torch.equal(out["final_boards"], ro.final_boards) and torch.equal(out["lengths"], la)
        alive = torch.arange(ro.t_steps, device="cuda").unsqueeze(1) < la.long().unsqueeze(0)
        want_scores = (ro.rewards.clamp(min=0) * alive).sum(dim=0).to(torch.int32)
        assert torch.equal(out["scores"], want_scores)
        s = out["summary"]
        assert s["episodes"] == batch and s["env_steps"] == ro.env_steps and s["loop_steps"] == ro.t_steps
        assert sum(s["max_tile_hist"].values()) == batch
        assert (rec.key == ev.key).all()
    with pytest.raises(ValueError):
        g2048.BatchRunner(init_seed=1, act_fn=g2048.act_drul).run_eval_batch(4)
