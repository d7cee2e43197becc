from g2048.actions import act_drul, act_randomly  # noqa: F401
