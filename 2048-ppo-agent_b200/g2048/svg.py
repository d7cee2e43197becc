"""SVG animation of a batch of 2048 games in the exact format ``pgx.save_svg_animation`` produces for the
reference (``run/viz_naive_strategies.py:115``, ``run/viz_ppo_agent.py:230``).

The layout rules were read off the reference's own artefacts (``assets/2048_drul_actions.svg`` and
``assets/2048_random_actions.svg``); for the runs behind those two files this writer reproduces them
byte for byte (``tests/test_svg.py`` checks the SHA-256).  Host-side string formatting only: boards are
fetched from the device once.  Pinned by the artefacts: 4 envs in a 2 x 2 grid, tiles up to 256.
Extrapolated: other batch sizes (square-ish grid), fill / text colours of tiles above 256.
"""
from __future__ import annotations

import math

import numpy as np

from . import engine as E


def _boards_of(frames) -> np.ndarray:
    """list of State / (B,16) arrays / (T,B,16) array -> uint8 (T, B, 16) exponents."""
    out = []
    for f in frames:
        if hasattr(f, "boards"):
            out.append(E.boards_numpy(f.boards))
        else:
            out.append(np.asarray(f, dtype=np.uint8).reshape(-1, 16))
    return np.stack(out)


def _fill(e: int) -> str:
    g = max(0, 242 - 22 * e)  # f2 (empty), dc (2), c6 (4), ... 42 (256)
    return f"#{g:02x}{g:02x}{g:02x}"


def _text_fill(e: int) -> str:
    if e <= 6:
        return "black"
    g = min(255, 145 + 10 * e)  # d7 (128), e1 (256)
    return f"#{g:02x}{g:02x}{g:02x}"


_FONT_SIZE = 18


def _text_x(col: int, digits: int) -> float:
    # centred for a 0.6 em wide Courier glyph; written like this it reproduces the reference files' float
    # formatting (e.g. 14.200000000000001 for a two-digit tile in column 0)
    return col * 50 + 25 - _FONT_SIZE * 0.3 * digits


def _board_group(board: np.ndarray, ox: float, oy: float) -> str:
    parts = [f'<g transform="translate({ox},{oy})">']
    for i in range(16):
        r, c = divmod(i, 4)
        e = int(board[i])
        parts.append(
            f'<rect fill="{_fill(e)}" height="46" rx="3px" ry="3px" stroke="black" stroke-width="0.5px" width="46" '
            f'x="{2 + 50 * c}" y="{2 + 50 * r}" />'
        )
        if e:
            value = str(1 << e)
            parts.append(
                f'<text fill="{_text_fill(e)}" font-family="Courier" font-size="{_FONT_SIZE}px" font-weight="bold" '
                f'x="{_text_x(c, len(value))}" y="{50 * r + 30.0}">{value}</text>'
            )
    parts.append("</g>")
    return "".join(parts)


def svg_animation(frames, frame_duration_seconds: float = 0.5) -> str:
    """The SVG document as a string.  frames: list of States (as run_actions_batch returns) or boards."""
    boards = _boards_of(frames)
    t_steps, batch, _ = boards.shape
    cols = math.ceil(math.sqrt(batch))
    rows = math.ceil(batch / cols)
    width, height = 250.0 * cols, 250.0 * rows
    total = t_steps * frame_duration_seconds
    pct = 100.0 / t_steps
    css = [f".frame{{visibility:hidden; animation:{total}s linear _k infinite;}}",
           f"@keyframes _k{{0%,{pct}%{{visibility:visible}}{pct * 1.000001}%,100%{{visibility:hidden}}}}"]
    for t in range(t_steps):
        css.append(f"#_fr{t:x}{{animation-delay:{t * frame_duration_seconds}s}}")
    doc = ['<?xml version="1.0" encoding="utf-8" ?>\n',
           f'<svg baseProfile="full" height="{height}" version="1.1" width="{width}" xmlns="http://www.w3.org/2000/svg" '
           'xmlns:ev="http://www.w3.org/2001/xml-events" xmlns:xlink="http://www.w3.org/1999/xlink">',
           '<defs><style type="text/css"><![CDATA[', "".join(css), "]]></style></defs>",
           '<rect fill="white" height="200" width="200" x="0" y="0" />' * batch]
    for t in range(t_steps):
        doc.append(f'<g class="frame" id="_fr{t:x}" transform="scale(1.0)">'
                   f'<rect fill="white" height="{int(height)}" width="{int(width)}" x="0" y="0" />')
        for b in range(batch):
            r, c = divmod(b, cols)
            doc.append(_board_group(boards[t, b], 25.0 + 250 * c, 25.0 + 250 * r))
            doc.append(f'<rect fill="none" height="250" stroke="gray" width="250" x="{250 * c}" y="{250 * r}" />')
        doc.append("</g>")
    doc.append("</svg>")
    return "".join(doc)


def save_svg_animation(frames, filename: str, frame_duration_seconds: float = 0.5) -> None:
    """pgx.save_svg_animation(states, filename, frame_duration_seconds=0.5) for 2048."""
    with open(filename, "w", encoding="utf-8") as f:
        f.write(svg_animation(frames, frame_duration_seconds))
