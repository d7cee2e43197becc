from .data_loader import DevicePPOBatches, PPODataset, compute_gae, create_ppo_dataloader
from .rollout_buffer import RolloutBuffer
from .torch_action_wrapper import TorchActionFunction

__all__ = ["DevicePPOBatches", "PPODataset", "RolloutBuffer", "TorchActionFunction", "compute_gae", "create_ppo_dataloader"]
