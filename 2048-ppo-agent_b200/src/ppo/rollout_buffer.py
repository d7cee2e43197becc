from g2048.ppo.rollout_buffer import RolloutBuffer  # noqa: F401
