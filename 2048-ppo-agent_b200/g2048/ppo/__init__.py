from .board_embedding import BoardEmbedding, embed_boards, forward_from_boards
from .collect import collect_rollouts
from .data_loader import DevicePPOBatches, PPODataset, compute_gae, create_ppo_dataloader
from .rollout_buffer import RolloutBuffer
from .torch_action_wrapper import TorchActionFunction
from .train_loop import PPOIterationLoop

__all__ = ["BoardEmbedding", "DevicePPOBatches", "PPODataset", "PPOIterationLoop", "RolloutBuffer", "TorchActionFunction", "compute_gae",
           "collect_rollouts", "create_ppo_dataloader", "embed_boards", "forward_from_boards"]
