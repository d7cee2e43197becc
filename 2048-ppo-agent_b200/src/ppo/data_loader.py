from g2048.ppo.data_loader import PPODataset, create_ppo_dataloader  # noqa: F401
