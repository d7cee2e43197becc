"""A batch of env states with the field names of pgx.State (v2 API).

The state itself is two device arrays -- the bitboards and one status byte per env -- plus the
rewards of the step that produced it.  The Pgx-shaped views (one-hot observation, bool mask,
terminated/truncated) are materialised on demand by libg2048 kernels.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import torch

from . import engine as E


@dataclass
class State:
    boards: torch.Tensor  # (B,) int64 bitboards
    status: torch.Tensor  # (B,) uint8
    reward: torch.Tensor  # (B,) float32
    _cache: dict = field(default_factory=dict, repr=False)

    @property
    def observation(self) -> torch.Tensor:
        """(B, 4, 4, 31) bool one-hot, empty cell -> channel 0 (pgx observe)."""
        if "obs" not in self._cache:
            self._cache["obs"] = E.expand_obs(self.boards, torch.bool).view(-1, 4, 4, 31)
        return self._cache["obs"]

    def _unpack(self):
        if "mask" not in self._cache:
            self._cache["mask"], self._cache["term"] = E.unpack_status(self.status)
        return self._cache["mask"], self._cache["term"]

    @property
    def legal_action_mask(self) -> torch.Tensor:
        return self._unpack()[0]

    @property
    def terminated(self) -> torch.Tensor:
        return self._unpack()[1]

    @property
    def truncated(self) -> torch.Tensor:
        return torch.zeros_like(self._unpack()[1])

    @property
    def rewards(self) -> torch.Tensor:
        """(B, 1) float32 like pgx State.rewards for a one-player game."""
        return self.reward.view(-1, 1)

    @property
    def _board(self) -> torch.Tensor:
        """(B, 16) int32 exponents (pgx keeps this as State._board)."""
        return torch.from_numpy(E.boards_numpy(self.boards).astype("int32"))

    def __len__(self) -> int:
        return self.boards.shape[0]
