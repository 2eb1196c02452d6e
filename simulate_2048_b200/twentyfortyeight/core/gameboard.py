"""Drop-in for ``twentyfortyeight/core/gameboard.py``.

Same signatures and conventions as the reference, including its quirks: ``slide_and_merge``
returns (score, board) with the score FIRST; ``merge_column`` returns a trimmed array;
``fill_cells`` mutates and returns its argument; ``next_state`` returns the very same object
and reward 0 for an illegal move.  All rules run on the device through the C ABI.
"""

from __future__ import annotations

import numpy as np
import torch
from numpy import ndarray

from ... import ops
from ..._cabi import FLAG_MOVED, FLAG_OVERFLOW
from ._device import SpawnStream, check_board, device, to_numpy, to_packed

# gameboard.py:13-18
TILE_SPAWN_PROBS: dict[int, float] = {2: 0.9, 4: 0.1}
_TILE_VALUES = [2, 4]
_TILE_PROBS = [0.9, 0.1]

# gameboard.py:20 — module-level generator (unseeded); rebind or .reseed() for determinism
_GENERATOR = SpawnStream()


def _overflow(flags) -> None:
    if int(flags) & FLAG_OVERFLOW:
        raise OverflowError('a 32768+32768 merge is not representable in the packed board format')


def _latent(state: ndarray, action: int):
    packed = to_packed(state)
    out = ops.expand(packed, children='none')
    _overflow(out.mask.item())
    a = int(action) % 4
    return out.after[:, a].contiguous(), int(out.rewards[0, a].item())


def merge_column(column: ndarray) -> tuple[int, ndarray]:
    """gameboard.py:23 — merge a 1-D line toward index 0; returns (score, trimmed line)."""
    col = np.asarray(column)
    if col.ndim != 1 or col.size > 4:
        raise NotImplementedError('only lines of up to 4 cells are implemented on the device path')
    board = np.zeros((4, 4), dtype=np.int64)
    board[0, : col.size] = col
    after, score = _latent(board, 0)
    row = to_numpy(after, board)[0]
    return score, row[row != 0].astype(col.dtype)


def slide_and_merge(board: ndarray) -> tuple[float, ndarray]:
    """gameboard.py:72 — slide left; returns (score, board): score FIRST."""
    b = check_board(board)
    after, score = _latent(b, 0)
    return float(score), to_numpy(after, b)


def latent_state(state: ndarray, action: int) -> tuple[ndarray, float]:
    """gameboard.py:105 — afterstate without spawn, and the reward."""
    s = check_board(state)
    after, score = _latent(s, action)
    return to_numpy(after, s), float(score)


def after_state(state: ndarray) -> list[tuple[ndarray, float]]:
    """gameboard.py:132 — every chance child (cell-major, value 2 then 4) with its probability."""
    s = check_board(state)
    packed = to_packed(s)
    children, num_empty, _ = ops.after_state(packed)
    k = int(num_empty.item())
    if k == 0:
        return [(s, 1.0)]  # gameboard.py:158-159: the state itself
    boards = to_numpy(children.reshape(-1)[: 2 * k].contiguous(), np.zeros((2 * k, 4, 4), dtype=s.dtype))
    return [(boards[j], TILE_SPAWN_PROBS[2 if j % 2 == 0 else 4] / k) for j in range(2 * k)]


def after_state_lazy(state: ndarray) -> tuple[ndarray, list[tuple[int, int]], int]:
    """gameboard.py:174 — (state, empty cells row-major, count) without building the children."""
    s = check_board(state)
    _, num_empty, cells = ops.after_state(to_packed(s), want_children=False)
    mask = int(cells.item()) & 0xFFFF
    return state, [(c // 4, c % 4) for c in range(16) if (mask >> c) & 1], int(num_empty.item())


def generate_outcome(state: ndarray, cell: tuple[int, int], value: int, num_empty: int) -> tuple[ndarray, float]:
    """gameboard.py:205 — one chance child on demand."""
    if num_empty <= 0:
        raise ValueError(f'num_empty must be > 0, got {num_empty}')
    s = check_board(state)
    byte = 0x80 | (int(cell[0]) * 4 + int(cell[1])) | (0x10 if value == 4 else 0)
    spawn = torch.tensor([byte], dtype=torch.uint8, device=device())
    new = ops.spawn(to_packed(s), spawn)
    return to_numpy(new, s), TILE_SPAWN_PROBS[value] / num_empty


def fill_cells(state: ndarray, number_tile: int, seed: int | None = None) -> ndarray:
    """gameboard.py:247 — add up to `number_tile` tiles IN PLACE; returns the same array."""
    s = check_board(state)
    packed = to_packed(s)
    for t in range(int(number_tile)):
        if seed is not None:
            packed = ops.spawn(packed, seed=int(seed), step=t, tag=1)
        else:
            packed = ops.spawn(packed, seed=_GENERATOR.key, step=_GENERATOR.next(), tag=1)
    state[...] = to_numpy(packed, s)
    return state


def next_state(state: ndarray, action: int, seed: int | None = None) -> tuple[ndarray, float]:
    """gameboard.py:291 — full transition; an illegal move returns (the same object, 0)."""
    s = check_board(state)
    act = torch.tensor([int(action) % 4], dtype=torch.uint8, device=device())
    if seed is not None:
        out = ops.step(to_packed(s), act, seed=int(seed))
    else:
        out = ops.step(to_packed(s), act, seed=_GENERATOR.key, step_base=_GENERATOR.next())
    flags = int(out.flags.item())
    _overflow(flags)
    if not flags & FLAG_MOVED:
        return state, 0  # gameboard.py:325
    return to_numpy(out.boards, s), float(out.rewards.item())


def is_done(state: ndarray) -> bool:
    """gameboard.py:328 — full board with no equal neighbours.  (An empty board is NOT done in
    the NumPy engine, unlike core.py:305.)"""
    s = check_board(state)
    flags = int(ops.legal_flags(to_packed(s)).item())
    return bool((flags & 0xF) == 0 and np.any(s != 0))
