from g2048.actions.act_randomly import act_randomly  # noqa: F401
