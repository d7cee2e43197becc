"""The ctypes binding printed in INTEGRATION.md is executed as it stands (only the library path is filled in): what
a maintainer would paste next to src/runs/batch_runner.py must load the library, play batches that equal the oracle,
and return the reference's GAE."""
import re
from pathlib import Path

import numpy as np
import pytest

from oracle import c_oracle as CO

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _stub_namespace():
    text = (ROOT / "INTEGRATION.md").read_text()
    blocks = re.findall(r"```python\n(.*?)```", text, flags=re.S)
    stub = next(b for b in blocks if "src/runs/_g2048.py" in b)
    stub = stub.replace("/path/to/repo/2048-ppo-agent_b200/libg2048.so", str(ROOT / "2048-ppo-agent_b200" / "libg2048.so"))
    ns = {}
    exec(compile(stub, "INTEGRATION.md:_g2048.py", "exec"), ns)
    return ns


def test_documented_binding_plays_and_computes_gae(golden_ppo):
    ns = _stub_namespace()
    for policy in (ns["POLICY_RANDOM"], ns["POLICY_DRUL"]):
        exps, lengths, scores, stats = ns["play_to_termination"](7, 3000, policy=policy)
        want = CO.play(7, 3000, policy, 1, max_steps=2048)
        np.testing.assert_array_equal(exps, want["final_boards"].astype(np.int32))
        np.testing.assert_array_equal(lengths, want["lengths"])
        np.testing.assert_array_equal(scores, want["scores"])
        assert int(stats[0]) == 3000 and int(stats[1]) == int(want["lengths"].sum())
    g = golden_ppo
    tag = next(k[len("gae_"):-len("_rewards")] for k in g.files if k.startswith("gae_") and k.endswith("_rewards"))
    gamma, lam = g[f"gae_{tag}_params"]
    adv, ret = ns["gae"](g[f"gae_{tag}_rewards"], g[f"gae_{tag}_values"], g[f"gae_{tag}_dones"], float(gamma), float(lam), normalize=False)
    np.testing.assert_array_equal(adv, g[f"gae_{tag}_adv"])
    np.testing.assert_array_equal(ret, g[f"gae_{tag}_ret"])
    adv_n, _ = ns["gae"](g[f"gae_{tag}_rewards"], g[f"gae_{tag}_values"], g[f"gae_{tag}_dones"], float(gamma), float(lam))
    np.testing.assert_allclose(adv_n, g[f"gae_{tag}_adv_norm"], rtol=1e-5, atol=1e-6)
