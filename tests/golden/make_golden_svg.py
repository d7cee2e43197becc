"""Decode the reference's golden trajectories into small fixtures.

Run in the build container only (it reads /root/reference, which does not exist on the
GPU box):

    python tests/golden/make_golden_svg.py

Source artefacts (produced by the real Pgx/JAX path):
  assets/2048_drul_actions.svg    <- run_actions_batch(0, 4, act_drul)
                                     (notebooks/explore_naive_strategies.ipynb cell 10)
  assets/2048_random_actions.svg  <- run_actions_batch(0, 4, act_randomly)  (cell 5)

Each SVG holds one <g class="frame" id="_frN"> per loop step; a frame holds four
<g transform="translate(x,y)"> groups (env order (25,25) (275,25) (25,275) (275,275)),
each with 16 <rect x= y=> cells (cell = (y//50, x//50)) and a <text> with the tile value
after the rect of every non-empty cell.  Frame t is the state AFTER step t+1.

Output: tests/golden/svg_trajectories.npz with
  drul_boards   uint8 (285, 4, 16)  exponents (0 = empty, e = tile 2**e)
  random_boards uint8 (123, 4, 16)
"""
import re
import sys
from pathlib import Path

import numpy as np

REF = Path("/root/reference/assets")
OUT = Path(__file__).resolve().parent / "svg_trajectories.npz"

ENV_ORIGINS = [(25.0, 25.0), (275.0, 25.0), (25.0, 275.0), (275.0, 275.0)]


def decode(path: Path) -> np.ndarray:
    text = path.read_text()
    frames = re.split(r'<g class="frame" id="_fr[0-9a-f]+"', text)[1:]
    out = np.zeros((len(frames), 4, 16), dtype=np.uint8)
    token = re.compile(
        r'<g transform="translate\(([\d.]+),([\d.]+)\)">'
        r'|<rect fill="#[0-9a-f]+" height="46"[^>]*? x="(\d+)" y="(\d+)" />'
        r"|<text [^>]*>(\d+)</text>"
    )
    for f, frame in enumerate(frames):
        env = -1
        cell = None
        seen = 0
        for m in token.finditer(frame):
            if m.group(1) is not None:
                env = ENV_ORIGINS.index((float(m.group(1)), float(m.group(2))))
                cell = None
            elif m.group(3) is not None:
                x, y = int(m.group(3)), int(m.group(4))
                cell = (y // 50) * 4 + (x // 50)
                seen += 1
            else:
                value = int(m.group(5))
                exponent = value.bit_length() - 1
                assert 1 << exponent == value and cell is not None and env >= 0
                out[f, env, cell] = exponent
        assert seen == 64, (f, seen)
    return out


def main() -> int:
    drul = decode(REF / "2048_drul_actions.svg")
    rand = decode(REF / "2048_random_actions.svg")
    assert drul.shape == (285, 4, 16), drul.shape
    assert rand.shape == (123, 4, 16), rand.shape
    np.savez_compressed(OUT, drul_boards=drul, random_boards=rand)
    # SHA-256 of the two files themselves: g2048.svg must regenerate them byte for byte (tests/test_svg.py)
    import hashlib
    import json

    hashes = {"_source": "sha256 of assets/2048_{drul,random}_actions.svg of the reference repo (tests/golden/make_golden_svg.py)"}
    for name in ("drul", "random"):
        raw = (REF / f"2048_{name}_actions.svg").read_bytes()
        hashes[name] = {"sha256": hashlib.sha256(raw).hexdigest(), "bytes": len(raw)}
    (OUT.parent / "svg_sha256.json").write_text(json.dumps(hashes, indent=1))
    print("wrote", OUT, drul.shape, rand.shape)
    print("drul final env0:\n", drul[-1, 0].reshape(4, 4))
    print("random final env3:\n", rand[-1, 3].reshape(4, 4))
    return 0


if __name__ == "__main__":
    sys.exit(main())
