"""Policy-network action function -- the reference's src/ppo/torch_action_wrapper.py:10-104.

The reference traces the PyTorch agent into JAX (torch2jax) so that the whole per-step call can
be vmapped inside XLA.  Here the network simply stays a PyTorch module (cuBLAS GEMMs, outside
the product path) and everything after it -- mask, clip, categorical draw with jax-compatible
Threefry bits, log-prob -- is the ``g2048_sample_logits`` kernel, or ``g2048_policy_step`` when
``BatchRunner`` fuses it with the env step and the record write.
"""
from __future__ import annotations

import torch

from .. import _native as N
from .. import engine as E
from ..actions._common import prepare
from ..env_definitions import BOARD_FLAT_DIM, OBS_DIM


class TorchActionFunction:
    def __init__(self, agent, use_mask: bool = False, sample_actions: bool = True,
                 device: torch.device = torch.device("cpu"), rng_mode=None, obs_dtype=torch.float32,
                 autocast_dtype=None):
        """
        agent          : module with forward(obs (B,16,31) float, mask|None) -> (logits (B,4), values (B,1))
        use_mask       : apply PPOAgent's mask rule `logits - 1e8 * (1 - mask)` (ppo_agent.py:117-121)
        sample_actions : categorical sample, else argmax
        device         : where the network runs; the sampling kernels always run on the CUDA device
        autocast_dtype : e.g. torch.bfloat16 -- run the network forward under torch.autocast (tensor-core GEMMs);
                         logits and values are converted back to float32 before the sampling kernel, which
                         always computes mask rule, log-softmax and draws in float32
        """
        self.cuda_device = N.require_cuda()
        self.device = torch.device(device)
        self.agent = agent.to(self.device).eval()
        self.use_mask = use_mask
        self.sample_actions = sample_actions
        self.rng_mode = E.resolve_rng_mode(rng_mode)
        self.obs_dtype = obs_dtype
        self.autocast_dtype = autocast_dtype

    @torch.no_grad()
    def forward_logits(self, obs: torch.Tensor):
        """obs (B,16,31) one-hot on the CUDA device -> raw logits (B,4) f32, values (B,) f32 (CUDA).
        The mask is NOT applied here: the sampling kernel applies the same formula in fp32."""
        x = obs.view(-1, BOARD_FLAT_DIM, OBS_DIM)
        if x.device != self.device:
            x = x.to(self.device)
        x = x.float() if x.dtype != torch.float32 and self.obs_dtype == torch.float32 else x
        if self.autocast_dtype is not None:
            with torch.autocast(self.device.type, dtype=self.autocast_dtype):
                logits, values = self.agent(x, None)
        else:
            logits, values = self.agent(x, None)
        logits = logits.float().to(self.cuda_device).contiguous()
        values = values.float().reshape(-1).to(self.cuda_device).contiguous()
        return logits, values

    def __call__(self, rng_key, obs, mask):
        """(key, obs (4,4,31), mask (4,)) -> (action, log_prob, value); a leading batch axis is accepted."""
        keys, status, batched = prepare(rng_key, obs, mask)
        obs_t = torch.as_tensor(obs).to(self.cuda_device).float().reshape(-1, BOARD_FLAT_DIM, OBS_DIM)
        logits, values = self.forward_logits(obs_t)
        actions, log_probs, _ = E.sample_logits(logits, status, self.use_mask, self.sample_actions, keys, 0, 0,
                                                self.rng_mode)
        if not batched:
            return actions[0], log_probs[0], values[0]
        return actions, log_probs, values
