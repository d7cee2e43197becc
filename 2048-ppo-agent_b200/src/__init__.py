"""Drop-in import paths of the reference (``src.env_definitions``, ``src.actions``, ``src.runs``,
``src.stats``, ``src.ppo``): put ``2048-ppo-agent_b200/`` on sys.path instead of the reference's
repo root and the same imports resolve to the B200 engine (package ``g2048``)."""

# Modules this package does not provide -- the learner (src/ppo/ppo_agent.py, transformer_encoder.py, ppo_trainer.py),
# src/optim -- can be taken from a checkout of the reference: with G2048_REFERENCE_ROOT=/path/to/2048-ppo-agent the
# reference's src/ directory is searched AFTER this one, so `from src.ppo.ppo_trainer import PPOTrainer` loads the
# reference's trainer while its `..runs.batch_runner`, `.rollout_buffer`, `.data_loader` and `.torch_action_wrapper`
# imports resolve to the B200 engine.
import os as _os

_ref = _os.environ.get("G2048_REFERENCE_ROOT")
if _ref and _os.path.isdir(_os.path.join(_ref, "src")):
    __path__.append(_os.path.join(_ref, "src"))
