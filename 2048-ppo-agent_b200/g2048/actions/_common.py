from __future__ import annotations

import numpy as np
import torch

from .. import _native as N
from ..env_definitions import ACTION_DIM, BOARD_DIM, OBS_DIM
from ..keys import as_key_tensor


def prepare(rng_key, obs, mask):
    """Shape checks of the reference (obs (4,4,31), mask (4,)) lifted to an optional batch axis.

    Returns (keys (n,2) int32 cuda, status (n,) uint8 cuda, batched flag).
    """
    dev = N.require_cuda()
    mask_t = torch.as_tensor(np.asarray(mask) if not isinstance(mask, torch.Tensor) else mask)
    batched = mask_t.dim() == 2
    assert mask_t.shape[-1] == ACTION_DIM and mask_t.dim() in (1, 2)
    if obs is not None:
        shape = tuple(obs.shape)
        assert shape[-3:] == (*BOARD_DIM, OBS_DIM), shape
        assert len(shape) == (4 if batched else 3), shape
    mask_t = mask_t.reshape(-1, ACTION_DIM).to(dev).to(torch.uint8)
    weights = torch.tensor([1, 2, 4, 8], dtype=torch.uint8, device=dev)
    status = (mask_t * weights).sum(dim=1).to(torch.uint8).contiguous()  # pack the caller's bools
    keys = as_key_tensor(rng_key, dev).reshape(-1, 2) if rng_key is not None else None
    if keys is not None:
        assert keys.shape[0] == status.shape[0], "one key per env"
    return keys, status, batched
