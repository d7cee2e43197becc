# The learner (PPOAgent, TransformerEncoder, PPOTrainer) is outside the hot path and stays the
# reference's own PyTorch code; only the rollout-side pieces are provided here.
from g2048.ppo import PPODataset, RolloutBuffer, TorchActionFunction, create_ppo_dataloader  # noqa: F401
