from .running_stats_vec import RunningStatsVec

__all__ = ["RunningStatsVec"]
