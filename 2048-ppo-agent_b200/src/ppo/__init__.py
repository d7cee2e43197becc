# The learner (PPOAgent, TransformerEncoder, PPOTrainer) is outside the hot path and stays the
# reference's own PyTorch code; only the rollout-side pieces are provided here.
from g2048.ppo import PPODataset, RolloutBuffer, TorchActionFunction, create_ppo_dataloader  # noqa: F401

# the reference's own modules of this sub-package (ppo_agent, transformer_encoder, ppo_trainer) are found behind ours
# when G2048_REFERENCE_ROOT is set (see src/__init__.py)
import os as _os

_ref = _os.environ.get("G2048_REFERENCE_ROOT")
if _ref and _os.path.isdir(_os.path.join(_ref, "src", "ppo")):
    __path__.append(_os.path.join(_ref, "src", "ppo"))
