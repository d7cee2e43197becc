from g2048.actions.act_drul import act_drul  # noqa: F401
