import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
PKG = ROOT / "2048-ppo-agent_b200"
for p in (str(ROOT), str(PKG)):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def golden_svg():
    import numpy as np

    return np.load(GOLDEN / "svg_trajectories.npz")


@pytest.fixture(scope="session")
def golden_ppo():
    import numpy as np

    return np.load(GOLDEN / "ppo_reference.npz")


@pytest.fixture(scope="session")
def golden_hist():
    import json

    return json.loads((GOLDEN / "histograms.json").read_text())
