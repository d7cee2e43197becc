"""Functional runner -- the reference's src/runs/run_actions_batch.py:10-57."""
from typing import Callable

from ..state import State
from .batch_runner import BatchRunner


def run_actions_batch(init_seed: int, batch_size: int, act_fn: Callable, rng_mode=None) -> list[State]:
    """Run `batch_size` envs to termination; returns the list of post-step States.

    Same key chain as ``BatchRunner`` (run_actions_batch.py:41-55); unlike
    ``BatchRunner.run_rollout_batch`` the init state is NOT part of the list (:46-57).
    """
    runner = BatchRunner(init_seed=init_seed, act_fn=act_fn, rng_mode=rng_mode)
    return runner.run_rollout_batch(batch_size)[1:]
