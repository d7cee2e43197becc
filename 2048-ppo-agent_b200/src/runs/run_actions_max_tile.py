from g2048.runs.run_actions_max_tile import run_actions_max_tile  # noqa: F401
