from g2048.ppo.torch_action_wrapper import TorchActionFunction  # noqa: F401
