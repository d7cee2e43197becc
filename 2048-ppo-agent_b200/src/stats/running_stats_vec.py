from g2048.stats.running_stats_vec import RunningStatsVec  # noqa: F401
