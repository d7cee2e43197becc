from g2048.stats import RunningStatsVec  # noqa: F401
