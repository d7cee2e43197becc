"""Packed boards straight into the policy's input embedding.

The reference's agent starts with ``nn.Linear(31, d_model, bias=False)`` applied to the float one-hot
observation ``(B, 16, 31)`` (src/ppo/ppo_agent.py:60,108).  On a one-hot row that product is a row of the
transposed weight, so the observation (1 984 bytes per board in float32) never has to be written or read:
``embed_boards(weight, boards)`` returns the same ``(B, 16, d_model)`` tensor from the 8-byte bitboards with
one ``g2048_embed_boards`` launch, and its backward is one ``g2048_embed_boards_grad`` launch.

``BoardEmbedding`` wraps an existing Linear (sharing its parameter) and ``forward_from_boards`` evaluates a
reference-shaped agent (attributes ``input_embedding``, ``transformer``, ``actor``, ``critic``, ``reduction``)
on packed boards; the network itself stays PyTorch.
"""
from __future__ import annotations

import torch

from .. import engine as E


class _EmbedBoards(torch.autograd.Function):
    @staticmethod
    def forward(ctx, weight: torch.Tensor, boards: torch.Tensor, indices, out_dtype):
        # (d_model, 31) -> (31, d_model) table in the dtype the Linear would have produced
        table = weight.detach().t().contiguous().to(out_dtype)
        ctx.save_for_backward(boards, indices if indices is not None else boards.new_empty(0))
        ctx.has_indices = indices is not None
        ctx.weight_dtype = weight.dtype
        return E.embed_boards(boards, table, indices)

    @staticmethod
    def backward(ctx, grad_out: torch.Tensor):
        boards, indices = ctx.saved_tensors
        if grad_out.dtype not in (torch.float32, torch.bfloat16):
            grad_out = grad_out.float()
        grad_table = E.embed_boards_grad(boards, grad_out.contiguous(), indices if ctx.has_indices else None)
        return grad_table.t().to(ctx.weight_dtype), None, None, None


def embed_boards(weight: torch.Tensor, boards: torch.Tensor, indices: torch.Tensor | None = None) -> torch.Tensor:
    """``F.linear(one_hot_observation(boards), weight)`` without the observation.

    weight: ``(d_model, 31)`` (an ``nn.Linear(31, d_model, bias=False).weight``), float32 or bfloat16, on the GPU.
    boards: int64 bitboards ``(B,)``; indices: optional int64 ``(M,)`` -- embed ``boards[indices]``.
    Under CUDA autocast the result has the autocast dtype, as the Linear's would.
    """
    if weight.dim() != 2 or weight.shape[1] != 31:
        raise ValueError(f"weight must be (d_model, 31), got {tuple(weight.shape)}")
    if boards.dtype != torch.int64:
        raise ValueError("boards must be int64 bitboards")
    out_dtype = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled("cuda") else weight.dtype
    if out_dtype not in (torch.float32, torch.bfloat16):
        raise ValueError(f"embed_boards computes in float32 or bfloat16, not {out_dtype}")
    return _EmbedBoards.apply(weight, boards.contiguous().view(-1), indices, out_dtype)


class BoardEmbedding(torch.nn.Module):
    """Drop-in for the agent's ``input_embedding`` that takes bitboards; shares the Linear's parameter."""

    def __init__(self, linear: torch.nn.Linear):
        super().__init__()
        if linear.bias is not None or linear.in_features != 31:
            raise ValueError("BoardEmbedding replaces Linear(31, d_model, bias=False)")
        self.weight = linear.weight

    def forward(self, boards: torch.Tensor, indices: torch.Tensor | None = None) -> torch.Tensor:
        return embed_boards(self.weight, boards, indices)


def forward_from_boards(agent, boards: torch.Tensor, action_mask: torch.Tensor | None = None,
                        indices: torch.Tensor | None = None):
    """``agent.forward(observations, action_mask)`` (src/ppo/ppo_agent.py:88-121) evaluated on packed boards."""
    embedded = embed_boards(agent.input_embedding.weight, boards, indices)
    features = agent.transformer(embedded, reduction=agent.reduction)
    logits = agent.actor(features)
    if action_mask is not None:
        logits = logits - (1e8 * (1 - action_mask.float()))
    return logits, agent.critic(features)
