"""A shortened run of tools/soak_parity.py under -m gpu: whole batches of the play kernel, every recorded step of a
window of the recording form, and every step of the fused policy step (auto-reset) against the oracle's C port, both
Threefry layouts, both built-in policies."""
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tools"))


@pytest.mark.parametrize("leg,envs", [("play", 40960), ("recorded", 40960), ("policy", 8192)])
def test_short_soak(leg, envs):
    import soak_parity

    lines = []
    total, bad = soak_parity.run(envs, 1, [leg], log=lines.append)
    assert bad == 0 and total > 0, "\n".join(lines)
