"""Shape constants of the 2048 environment under the reference's names (src/env_definitions.py:1-8)."""
BOARD_DIM = (4, 4)                               # rows, columns
BOARD_FLAT_DIM = BOARD_DIM[0] * BOARD_DIM[1]     # 16 cells, cell i = 4 * row + column = nibble i of the bitboard
OBS_DIM = 31                                     # one-hot channels per cell: tile exponent 0 (empty) .. 30
ACTION_DIM = 4                                   # 0 Left, 1 Up, 2 Right, 3 Down
