"""g2048 -- B200-native rollout engine for 2048 (env step + action selection + rollout records + GAE).

The compute lives in libg2048.so (hand-written sm_100a CUDA behind a C ABI, include/g2048.h);
this package is the thin Python host side that mirrors the reference's own API
(src/env_definitions.py, src/actions, src/runs, src/stats, src/ppo rollout pieces).
"""
from . import _native  # noqa: F401  loads libg2048.so; raises ImportError if it is missing
from . import engine  # noqa: F401

__all__ = ["engine"]
