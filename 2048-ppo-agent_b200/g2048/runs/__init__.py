from .batch_runner import BatchRunner, FlatRollout, PackedRollout
from .fixed_horizon import FixedHorizonRunner, FixedRollout
from .run_actions_batch import run_actions_batch
from .run_actions_max_tile import run_actions_max_tile

__all__ = ["BatchRunner", "FixedHorizonRunner", "FixedRollout", "FlatRollout", "PackedRollout", "run_actions_batch", "run_actions_max_tile"]
