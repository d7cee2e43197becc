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
