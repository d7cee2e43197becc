"""g2048_policy_step_obs = g2048_policy_step_at + g2048_expand_obs of the stepped boards in one launch: same state,
records and draws, and the observation tensor of the next forward pass (src/runs/batch_runner.py:117-136 with
src/ppo/torch_action_wrapper.py:84-102)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def E():
    from g2048 import engine

    return engine


def _records(steps, n):
    return (torch.zeros((steps, n), dtype=torch.int64, device="cuda"), torch.zeros((steps, n), dtype=torch.uint8, device="cuda"),
            torch.zeros((steps, n), dtype=torch.float32, device="cuda"), torch.zeros((steps, n), dtype=torch.float32, device="cuda"),
            torch.zeros((steps, n), dtype=torch.float32, device="cuda"))


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.bool])
@pytest.mark.parametrize("n,auto_reset", [(1, False), (33, True), (1000, False), (70001, True)])
def test_fused_step_equals_step_then_expand(E, mode, dtype, n, auto_reset):
    steps, batch, lo = 6, n + 11, 5
    gen = torch.Generator(device="cuda").manual_seed(n)
    subs = E.chain_advance(E.words_tensor([0, 77], "cuda"), mode, 1 + 2 * steps)
    boards_a, status_a = E.env_init(subs[0], batch, lo, n, mode)
    # start some envs close to the end so that terminations (and auto-resets) happen within the few steps
    full = torch.tensor(E.pack_boards([[1, 2, 1, 2, 2, 1, 2, 1, 1, 2, 1, 2, 2, 1, 0, 1]]), device="cuda")
    boards_a[::3] = full[0]
    status_a[::3] = 0x0F
    boards_b, status_b = boards_a.clone(), status_a.clone()
    rec_a, rec_b = _records(steps, n), _records(steps, n)
    acts_a = torch.zeros(n, dtype=torch.int32, device="cuda")
    acts_b = torch.zeros_like(acts_a)
    obs_b = torch.empty((n, 16, 31), dtype=dtype, device="cuda")
    counters = torch.zeros(2, dtype=torch.int64, device="cuda")
    step_index = torch.zeros((), dtype=torch.int32, device="cuda")
    done_before = int(((status_a & 0x10) != 0).sum())
    reward_sum = 0
    for t in range(steps):
        logits = torch.randn((n, 4), device="cuda", generator=gen)
        values = torch.randn(n, device="cuda", generator=gen)
        E.policy_step(boards_a, status_a, logits, values, True, True, auto_reset, subs[1 + 2 * t], subs[2 + 2 * t], batch, lo, mode,
                      rec_a[0][t], rec_a[1][t], rec_a[2][t], rec_a[3][t], rec_a[4][t], acts_a)
        if t % 2 == 0:  # explicit sub key pointer and record rows
            E.policy_step_obs(boards_b, status_b, logits, values, True, True, auto_reset, subs[1 + 2 * t:], None, batch, lo, mode,
                              obs_b, rec_b[0][t], rec_b[1][t], rec_b[2][t], rec_b[3][t], rec_b[4][t], acts_b, counters)
            step_index += 1
        else:  # step number from device memory, advanced by the kernel itself
            E.policy_step_obs(boards_b, status_b, logits, values, True, True, auto_reset, subs[1:], step_index, batch, lo, mode,
                              obs_b, *rec_b, acts_b, counters, advance_step=True)
        assert int(step_index) == t + 1
        assert torch.equal(boards_a, boards_b) and torch.equal(status_a, status_b) and torch.equal(acts_a, acts_b)
        assert torch.equal(obs_b, E.expand_obs(boards_a, dtype))
        reward_sum += int(rec_a[2][t].clamp(min=0).sum())
    for x, y in zip(rec_a, rec_b):
        assert torch.equal(x, y)
    dones = int(((rec_a[1] >> 6) & 1).sum()) if auto_reset else int(((status_a & 0x10) != 0).sum()) - done_before
    assert counters.tolist() == [dones, reward_sum]


def test_runner_paths_agree_with_the_fused_step():
    """eager (fused launch per step, one sync per chunk), CUDA graph, and live-env compaction: same packed rollouts."""
    import g2048

    class RowwiseAgent(torch.nn.Module):
        """A row's outputs depend on that row only AND are computed with a fixed operation order whatever the batch
        shape: per-cell table rows picked by the cell's exponent, added cell after cell with elementwise adds (a
        reduction kernel or a GEMM may split its sums differently for different batch sizes)."""

        def __init__(self):
            super().__init__()
            g = torch.Generator().manual_seed(3)
            self.w = torch.nn.Parameter(torch.randn(16, 31, 5, generator=g) * 0.5)

        def forward(self, obs, mask=None):
            idx = obs.reshape(obs.shape[0], 16, 31).argmax(dim=-1)  # (B, 16) exponents
            cells = torch.arange(16, device=obs.device)
            vals = self.w[cells.unsqueeze(0), idx]  # (B, 16, 5)
            out = vals[:, 0]
            for c in range(1, 16):
                out = out + vals[:, c]
            return out[:, :4], out[:, 4:5]

    def make(**kw):
        fn = g2048.TorchActionFunction(RowwiseAgent(), use_mask=True, device=torch.device("cuda"))
        return g2048.BatchRunner(init_seed=12, act_fn=fn, **kw)

    eager, graph, live = make(), make(cuda_graph=True), make(compact_live=True)
    for batch in (70, 300):
        a, b, c = eager.run_packed_batch(batch), graph.run_packed_batch(batch), live.run_packed_batch(batch)
        assert a.t_steps == b.t_steps == c.t_steps and a.env_steps == b.env_steps == c.env_steps
        la = a.lengths().long()
        alive = torch.arange(a.t_steps, device="cuda").unsqueeze(1) < la.unsqueeze(0)
        for name in ("boards", "meta", "rewards", "log_probs", "values"):
            assert torch.equal(getattr(a, name), getattr(b, name)), name
            assert torch.equal(getattr(a, name)[alive], getattr(c, name)[alive]), name
            assert bool((getattr(c, name)[~alive] == 0).all()), name
        assert torch.equal(a.final_boards, b.final_boards) and torch.equal(a.final_boards, c.final_boards)
        assert (eager.key == graph.key).all() and (eager.key == live.key).all()
    # the records of a run equal a step-by-step replay with the public one-step kernels
    ref = make()
    obs, actions, masks, log_probs, values, rewards, terms = ref.run_actions_batch(70)
    ro = make().run_packed_batch(70)
    assert actions.shape[1] == ro.t_steps
    assert (actions.T == (ro.meta & 3).cpu().numpy()).all() and (rewards.T == ro.rewards.cpu().numpy()).all()
