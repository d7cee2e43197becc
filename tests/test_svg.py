"""g2048.svg regenerates the reference's SVG animations byte for byte."""
import hashlib
import json

import numpy as np
import pytest

from conftest import GOLDEN


@pytest.fixture(scope="module")
def hashes():
    return json.loads((GOLDEN / "svg_sha256.json").read_text())


@pytest.mark.parametrize("name", ["drul", "random"])
def test_svg_writer_reproduces_reference_files_from_golden_boards(golden_svg, hashes, name):
    from g2048 import svg

    doc = svg.svg_animation(list(golden_svg[f"{name}_boards"])).encode("utf-8")
    assert len(doc) == hashes[name]["bytes"]
    assert hashlib.sha256(doc).hexdigest() == hashes[name]["sha256"]


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["drul", "random"])
def test_cuda_rollout_to_reference_svg_end_to_end(hashes, name, tmp_path):
    """notebooks/explore_naive_strategies.ipynb cells 5 / 10: run_actions_batch(0, 4, act_fn) then
    save_svg_animation -- kernels, host API and writer together give the reference's file, byte for byte."""
    import g2048
    from g2048 import svg

    fn = g2048.act_drul if name == "drul" else g2048.act_randomly
    states = g2048.run_actions_batch(0, 4, fn, rng_mode="original")
    path = tmp_path / f"2048_{name}_actions.svg"
    svg.save_svg_animation(states, str(path), frame_duration_seconds=0.5)
    raw = path.read_bytes()
    assert len(raw) == hashes[name]["bytes"] and hashlib.sha256(raw).hexdigest() == hashes[name]["sha256"]


def test_svg_writer_other_batch_sizes_are_well_formed(golden_svg):
    from g2048 import svg
    import xml.etree.ElementTree as ET

    boards = np.concatenate([golden_svg["random_boards"][:5], golden_svg["random_boards"][:5, :1]], axis=1)  # 5 envs
    boards[0, 0, 0] = 11  # a 2048 tile
    root = ET.fromstring(svg.svg_animation(list(boards)))
    frames = [g for g in root.iter("{http://www.w3.org/2000/svg}g") if g.get("class") == "frame"]
    assert len(frames) == 5 and root.get("width") == "750.0" and root.get("height") == "500.0"
