from g2048.runs.batch_runner import ENV_ID, BatchRunner  # noqa: F401
