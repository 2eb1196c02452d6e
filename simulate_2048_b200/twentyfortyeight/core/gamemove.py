"""Drop-in for ``twentyfortyeight/core/gamemove.py``: legal-move queries on a NumPy board,
computed by ``g2048_mask`` on the device."""

from __future__ import annotations

from numpy import ndarray

from ... import ops
from ._device import check_board, to_packed


def _mask_bits(state: ndarray) -> int:
    return int(ops.legal_flags(to_packed(check_board(state))).item()) & 0xF


def legal_actions_mask(state: ndarray) -> tuple[bool, bool, bool, bool]:
    """gamemove.py:45 — (left, up, right, down), True = legal."""
    m = _mask_bits(state)
    return bool(m & 1), bool(m & 2), bool(m & 4), bool(m & 8)


def illegal_actions(state: ndarray) -> list[int]:
    """gamemove.py:86."""
    m = _mask_bits(state)
    return [a for a in range(4) if not (m >> a) & 1]


def legal_actions(state: ndarray) -> list[int]:
    """gamemove.py:109."""
    m = _mask_bits(state)
    return [a for a in range(4) if (m >> a) & 1]


def _can_move_direction(board: ndarray, direction: int) -> bool:
    """gamemove.py:9 — direction 0 left, 1 up, 2 right, anything else down."""
    d = direction if direction in (0, 1, 2) else 3
    return bool((_mask_bits(board) >> d) & 1)


def can_move(board: ndarray) -> bool:
    """gamemove.py:132 — can anything move LEFT on this board."""
    return bool(_mask_bits(board) & 1)
