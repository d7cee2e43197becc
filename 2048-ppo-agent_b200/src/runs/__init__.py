from g2048.runs import BatchRunner, FixedHorizonRunner, run_actions_batch, run_actions_max_tile  # noqa: F401
