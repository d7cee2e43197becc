"""Max-tile statistics of a policy -- the reference's src/runs/run_actions_max_tile.py:10-71."""
import warnings
from typing import Callable

import numpy as np

from .. import engine as E
from ..stats.running_stats_vec import RunningStatsVec
from .batch_runner import BatchRunner


def run_actions_max_tile(init_seed: int, batch_size: int, num_envs: int, act_fn: Callable, rng_mode=None,
                         exact_reference_quirk: bool = True) -> RunningStatsVec:
    """Runs num_envs // batch_size batches and accumulates the max tile of every episode.

    For act_randomly / act_drul each batch is one persistent ``g2048_play`` launch -- no (B,T,16,31) observations are
    materialised for the argmax / 2** / max of run_actions_max_tile.py:61-67.

    exact_reference_quirk: the reference reads the last stored PRE-step observation (:64 with
    batch_runner.py:121,130), so for the env(s) that live until the last loop step the final merge
    and spawn are not counted.  True reproduces that -- ``g2048_replay_envs`` replays just those envs (usually one) up to
    their last step --; False uses each env's real final board.
    """
    if num_envs % batch_size != 0:
        warnings.warn(
            f"The number of environments ({num_envs}) is not divisible by the batch size ({batch_size}). "
            "The number of environments will be adjusted to be divisible by the batch size."
        )
    runner = BatchRunner(init_seed=init_seed, act_fn=act_fn, rng_mode=rng_mode)
    stats = RunningStatsVec()
    fused = getattr(act_fn, "policy_id", None) is not None
    for _ in range(num_envs // batch_size):
        if fused:
            out = runner.run_stats_batch(batch_size, per_env=True, last_stored_boards=exact_reference_quirk)
            final_boards = out["last_stored_boards"] if exact_reference_quirk else out["final_boards"]
        else:
            ro = runner.run_packed_batch(batch_size)
            # last stored observation = the board BEFORE the last loop step (frozen envs keep theirs)
            final_boards = ro.boards[ro.t_steps - 1] if exact_reference_quirk else ro.final_boards
        exps = E.boards_numpy(final_boards)
        max_tiles = (2 ** exps.astype(np.int64)).max(-1)
        stats.push(max_tiles.reshape(1, -1).astype(np.float64))
    return stats
