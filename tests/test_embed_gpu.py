"""Packed boards -> input embedding against the reference's own formulation:
Linear(31 -> d_model, bias=False) on the float one-hot observation (src/ppo/ppo_agent.py:60,108)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _boards(n, seed=0):
    """Random boards with exponents 0..15 (0 most common, like real positions) as int64 bitboards."""
    rng = np.random.default_rng(seed)
    cells = rng.integers(0, 16, (n, 16)) * (rng.random((n, 16)) < 0.7)
    from g2048 import engine as E
    return torch.from_numpy(E.pack_boards(cells).view(np.int64)).cuda(), cells


@pytest.mark.parametrize("entry", ["g2048_embed_boards", "g2048_embed_boards_bulk", "g2048_embed_boards_plain"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("n,d_model", [(1, 256), (777, 256), (5000, 64), (333, 8), (41, 1024)])
def test_forward_is_the_linear_on_one_hot_observations(entry, dtype, n, d_model):
    from g2048 import engine as E
    boards, cells = _boards(n, n)
    torch.manual_seed(d_model)
    weight = torch.randn(d_model, 31, device="cuda").to(dtype)
    got = E.embed_boards(boards, weight.t().contiguous(), entry=entry)
    assert got.shape == (n, 16, d_model) and got.dtype == dtype
    want = weight.t()[torch.from_numpy(cells).cuda().long()]  # row gather == one-hot @ W^T exactly
    assert torch.equal(got, want)
    if dtype == torch.float32:
        obs = E.expand_obs(boards, torch.float32)
        lin = torch.nn.functional.linear(obs.double(), weight.double()).float()  # exact: one non-zero term per sum
        assert torch.equal(got, lin)


def test_forward_with_indices_gathers_boards():
    from g2048 import engine as E
    boards, _ = _boards(4096, 3)
    table = torch.randn(31, 256, device="cuda")
    idx = torch.randperm(4096, device="cuda")[:1000].contiguous()
    assert torch.equal(E.embed_boards(boards, table, idx), E.embed_boards(boards[idx].contiguous(), table))


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("n,d_model", [(1, 256), (777, 256), (20000, 256), (3000, 64), (100, 1024), (50, 8)])
def test_table_gradient_matches_the_linear_backward(dtype, n, d_model):
    from g2048 import engine as E
    boards, cells = _boards(n, 7 * n)
    grad = torch.randn(n, 16, d_model, device="cuda").to(dtype)
    got = E.embed_boards_grad(boards, grad)
    assert got.shape == (31, d_model) and got.dtype == torch.float32
    onehot = torch.nn.functional.one_hot(torch.from_numpy(cells).cuda().long(), 31).double()  # (n,16,31)
    want = torch.einsum("ncr,ncd->rd", onehot, grad.double())
    scale = float(want.abs().max()) + 1.0
    assert float((got.double() - want).abs().max()) <= 2e-6 * scale * max(1.0, np.sqrt(n * 16 / 31))
    assert float(got[16:].abs().max()) == 0.0
    again = E.embed_boards_grad(boards, grad)
    assert torch.equal(got, again), "summation order must not depend on scheduling"


def test_gradient_with_indices():
    from g2048 import engine as E
    boards, _ = _boards(4096, 5)
    idx = torch.randint(0, 4096, (1500,), device="cuda")
    grad = torch.randn(1500, 16, 128, device="cuda")
    assert torch.equal(E.embed_boards_grad(boards, grad, idx), E.embed_boards_grad(boards[idx].contiguous(), grad))


def test_autograd_function_matches_linear_on_observations():
    from g2048 import engine as E
    from g2048.ppo import BoardEmbedding, embed_boards
    boards, _ = _boards(2048, 11)
    lin = torch.nn.Linear(31, 256, bias=False).cuda()
    emb = BoardEmbedding(lin)
    assert emb.weight is lin.weight
    head = torch.randn(256, device="cuda")
    out = emb(boards)
    (out * head).sum().backward()
    g_boards = lin.weight.grad.clone()
    lin.weight.grad = None
    ref = lin(E.expand_obs(boards, torch.float32))
    (ref * head).sum().backward()
    assert torch.equal(out, ref.detach()) or torch.allclose(out, ref.detach(), rtol=0, atol=0)
    assert torch.allclose(g_boards, lin.weight.grad, rtol=1e-5, atol=1e-3)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        a = embed_boards(lin.weight, boards)
        b = lin(E.expand_obs(boards, torch.float32))
    assert a.dtype == b.dtype == torch.bfloat16 and torch.equal(a, b)


def test_agent_forward_from_boards():
    from g2048 import engine as E
    from g2048.ppo import forward_from_boards

    class Encoder(torch.nn.Module):  # stands in for the reference's TransformerEncoder (outside the product path)
        def forward(self, x, reduction="mean"):
            return x.mean(dim=1) if reduction == "mean" else x[:, 0]

    class Agent(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.input_embedding = torch.nn.Linear(31, 64, bias=False)
            self.transformer = Encoder()
            self.actor = torch.nn.Linear(64, 4, bias=False)
            self.critic = torch.nn.Linear(64, 1, bias=False)
            self.reduction = "mean"

        def forward(self, observations, action_mask=None):  # src/ppo/ppo_agent.py:88-121
            f = self.transformer(self.input_embedding(observations), reduction=self.reduction)
            logits = self.actor(f)
            if action_mask is not None:
                logits = logits - (1e8 * (1 - action_mask.float()))
            return logits, self.critic(f)

    agent = Agent().cuda()
    boards, _ = _boards(512, 2)
    mask = torch.rand(512, 4, device="cuda") < 0.7
    l0, v0 = agent(E.expand_obs(boards, torch.float32), mask)
    l1, v1 = forward_from_boards(agent, boards, mask)
    assert torch.equal(l0, l1) and torch.equal(v0, v1)


def test_minibatches_can_carry_boards_instead_of_observations():
    import g2048
    runner = g2048.BatchRunner(init_seed=3, act_fn=g2048.act_randomly)
    buf = g2048.RolloutBuffer(31, 16, 4)
    buf.store_packed(runner.run_packed_batch(64))
    gen = torch.Generator(device="cuda"); gen.manual_seed(1)
    with_obs = g2048.DevicePPOBatches(buf.get_packed(), batch_size=128, generator=gen)
    gen2 = torch.Generator(device="cuda"); gen2.manual_seed(1)
    with_boards = g2048.DevicePPOBatches(buf.get_packed(), batch_size=128, obs_dtype=None, generator=gen2)
    from g2048 import engine as E
    seen = 0
    for a, b in zip(with_obs, with_boards):
        assert "observations" not in b and b["boards"].dtype == torch.int64
        assert torch.equal(a["observations"], E.expand_obs(b["boards"], torch.float32))
        for k in ("actions", "action_masks", "log_probs", "values", "advantages", "returns"):
            assert torch.equal(a[k], b[k])
        seen += 1
    assert seen > 0


def test_validation():
    from g2048 import engine as E
    from g2048.ppo import embed_boards
    boards, _ = _boards(4)
    with pytest.raises(ValueError):
        E.embed_boards(boards, torch.zeros(30, 256, device="cuda"))
    with pytest.raises(RuntimeError, match="multiple of 16"):
        E.embed_boards(boards, torch.zeros(31, 6, device="cuda"))
    with pytest.raises(ValueError):
        embed_boards(torch.zeros(256, 30, device="cuda"), boards)
    with pytest.raises(ValueError):
        E.embed_boards_grad(boards, torch.zeros(4, 16, 6, device="cuda"))


def test_reused_minibatch_buffers_hold_the_same_batches():
    import g2048
    runner = g2048.BatchRunner(init_seed=5, act_fn=g2048.act_randomly)
    buf = g2048.RolloutBuffer(31, 16, 4)
    buf.store_packed(runner.run_packed_batch(64))
    for obs_dtype in (torch.float32, None):
        gens = [torch.Generator(device="cuda") for _ in range(2)]
        for g in gens:
            g.manual_seed(2)
        fresh = g2048.DevicePPOBatches(buf.get_packed(), batch_size=100, drop_last=False, obs_dtype=obs_dtype, generator=gens[0])
        reused = g2048.DevicePPOBatches(buf.get_packed(), batch_size=100, drop_last=False, obs_dtype=obs_dtype, generator=gens[1],
                                        reuse_buffers=True)
        previous, ptrs = None, set()
        for a, b in zip(fresh, reused):
            assert a.keys() == b.keys()
            for k in a:
                assert torch.equal(a[k], b[k]), k
            if previous is not None and previous[1]["actions"].shape[0] == 100:  # the batch before stays valid
                for k in previous[0]:
                    assert torch.equal(previous[0][k], previous[1][k]), k
            previous = (a, b)
            ptrs.add(b["actions"].data_ptr())
        assert len(ptrs) <= 3  # two alternating sets (+ a fresh one for the short last batch)
