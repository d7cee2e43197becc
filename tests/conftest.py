import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
PKG = ROOT / "2048-ppo-agent_b200"
for p in (str(ROOT), str(PKG)):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = ROOT / "tests" / "golden"


LEGACY_LIB = ROOT / "tests" / "legacy" / "libg2048_legacy.so"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")
    # the first-generation kernels (g2048_*_v1) the current ones are tested against live in a test-only build of the
    # library (tests/legacy/g2048_legacy.h); __graft_entry__.build() compiles it, `make legacy` here as a fallback
    import subprocess

    if not LEGACY_LIB.exists():
        subprocess.run(["make", "-C", str(PKG / "csrc"), "legacy", "-j8"], check=False, capture_output=True)
    if LEGACY_LIB.exists():
        from g2048 import _native as N

        N.register_entry_points(LEGACY_LIB, N.LEGACY_SIGNATURES)


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
