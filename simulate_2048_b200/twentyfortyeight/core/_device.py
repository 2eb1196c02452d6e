"""Host <-> device glue shared by gameboard.py / gamemove.py (no game rules live here)."""

from __future__ import annotations

import os

import numpy as np
import torch

from ... import ops


def device() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError('simulate_2048_b200 needs a CUDA device (there is no CPU fallback)')
    return torch.device('cuda', torch.cuda.current_device())


def check_board(state: np.ndarray) -> np.ndarray:
    a = np.asarray(state)
    if a.shape != (4, 4):
        raise NotImplementedError('only the 4x4 board is implemented on the device path')
    return a


def to_packed(state: np.ndarray) -> torch.Tensor:
    """[4,4] (or [N,4,4]) tile values -> packed int64[N] on the device (g2048_pack)."""
    a = np.asarray(state)
    t = torch.as_tensor(np.ascontiguousarray(a, dtype=np.int32).reshape(-1, 16)).to(device())
    return ops.pack(t)


def to_numpy(packed: torch.Tensor, like: np.ndarray) -> np.ndarray:
    """packed int64[N] -> tile values with the dtype/shape of `like` (g2048_unpack)."""
    tiles = ops.unpack(packed).cpu().numpy()
    like = np.asarray(like)
    return tiles.reshape(like.shape).astype(like.dtype if like.dtype.kind in 'iu' else np.int64)


def pack_host(state: np.ndarray):
    """One [4,4] board of tile values -> its packed 64-bit word, on the host (a change of
    representation, not a game rule).  Returns None when some value is not 0 or a power of two in
    [2, 32768] — the caller then takes the device pack path, which reports the error."""
    word = 0
    for i, v in enumerate(np.asarray(state).reshape(-1).tolist()):
        v = int(v)
        if v:
            e = v.bit_length() - 1
            if v < 2 or v > 32768 or (1 << e) != v:
                return None
            word |= e << (4 * i)
    return word


def unpack_host(word: int, dtype=np.int64) -> np.ndarray:
    """Packed word -> [4,4] tile values (host)."""
    word = int(word)
    return np.array([(1 << ((word >> (4 * i)) & 0xF)) if (word >> (4 * i)) & 0xF else 0 for i in range(16)],
                    dtype=dtype).reshape(4, 4)


class SpawnStream:
    """Stand-in for the module-level ``_GENERATOR`` (gameboard.py:20): a Philox key plus a draw
    counter.  Unseeded at import, like the reference."""

    def __init__(self, seed=None):
        self.reseed(seed)

    def reseed(self, seed=None):
        self.key = int.from_bytes(os.urandom(8), 'little') if seed is None else int(seed) & (2**64 - 1)
        self.count = 0

    def next(self) -> int:
        c = self.count
        self.count += 1
        return c
