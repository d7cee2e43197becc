from g2048.runs.run_actions_batch import run_actions_batch  # noqa: F401
from g2048.state import State  # noqa: F401
