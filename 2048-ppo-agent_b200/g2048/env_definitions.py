"""Same constants as the reference's src/env_definitions.py:1-8."""
# Observation dimension
OBS_DIM = 31
# Board dimension
BOARD_DIM = (4, 4)
# Board flat dimension
BOARD_FLAT_DIM = 16
# Action dimension
ACTION_DIM = 4
