"""Batched ndarray entry points of the NumPy-engine mirror.

The reference engine takes ONE board per call (``gameboard.py:105-325``, ``gamemove.py:45``); its
callers loop in Python.  A drop-in that kept that shape would spend ~100 µs of host/device round
trip per board, so the mirror also offers the same functions over ``[N, 4, 4]`` arrays of tile
values: one host-to-device copy, one kernel, one copy back.  Semantics per board are exactly those
of the single-board functions of ``gameboard.py`` / ``gamemove.py`` in this package (and thus of
the reference); every rule runs in the CUDA kernels behind the C ABI.
"""

from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch

from ... import ops
from ..._cabi import FLAG_DONE, FLAG_MOVED, FLAG_OVERFLOW
from ._device import SpawnStream, device

_STREAM = SpawnStream()


def _boards(states: np.ndarray) -> Tuple[torch.Tensor, tuple, np.dtype]:
    a = np.asarray(states)
    if a.ndim < 2 or a.shape[-2:] != (4, 4):
        raise NotImplementedError('only 4x4 boards are implemented on the device path')
    tiles = torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32).reshape(-1, 16)).to(device(), non_blocking=True)
    return ops.pack(tiles), a.shape[:-2], (a.dtype if a.dtype.kind in 'iu' else np.dtype(np.int64))


def _tiles(packed: torch.Tensor, lead: tuple, dtype) -> np.ndarray:
    return ops.unpack(packed.reshape(-1)).cpu().numpy().reshape(*lead, 4, 4).astype(dtype, copy=False)


def _actions(actions, n: int) -> torch.Tensor:
    a = np.asarray(actions).reshape(-1)
    if a.size == 1 and n != 1:
        a = np.repeat(a, n)
    if a.size != n:
        raise ValueError('actions: one entry per board')
    return torch.from_numpy((a % 4).astype(np.uint8)).to(device(), non_blocking=True)


def _check_overflow(flags: torch.Tensor) -> None:
    if bool((flags & FLAG_OVERFLOW).any().item()):
        raise OverflowError('a 32768+32768 merge is not representable in the packed board format')


def latent_state_batch(states: np.ndarray, actions) -> Tuple[np.ndarray, np.ndarray]:
    """gameboard.py:105 over a batch: (afterstates [N,4,4], rewards float64[N]); no spawn."""
    packed, lead, dtype = _boards(states)
    n = packed.numel()
    ex = ops.expand(packed, children='none')
    _check_overflow(ex.mask)
    idx = _actions(actions, n).to(torch.int64).unsqueeze(1)
    after = ex.after.gather(1, idx).squeeze(1)
    rewards = ex.rewards.gather(1, idx).squeeze(1)
    return _tiles(after, lead, dtype), rewards.cpu().numpy().astype(np.float64).reshape(lead)


def next_state_batch(states: np.ndarray, actions, seed: Optional[int] = None,
                     spawns: Optional[np.ndarray] = None) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """gameboard.py:291 over a batch: (new states, rewards float64[N], moved bool[N]).  An illegal
    move returns the board unchanged with reward 0 (gameboard.py:325).  ``spawns`` (uint8[N] ABI spawn
    bytes) injects the tiles; otherwise the Philox stream keyed by (seed, board index) — or by the
    module stream and its draw counter when ``seed`` is None, as the reference's ``_GENERATOR``."""
    packed, lead, dtype = _boards(states)
    n = packed.numel()
    acts = _actions(actions, n)
    if spawns is not None:
        sp = torch.from_numpy(np.ascontiguousarray(spawns, dtype=np.uint8).reshape(-1)).to(device())
        out = ops.step(packed, acts, sp)
    elif seed is not None:
        out = ops.step(packed, acts, seed=int(seed))
    else:
        out = ops.step(packed, acts, seed=_STREAM.key, step_base=_STREAM.next())
    _check_overflow(out.flags)
    flags = out.flags.cpu().numpy()
    return (_tiles(out.boards, lead, dtype), out.rewards.cpu().numpy().astype(np.float64).reshape(lead),
            ((flags & FLAG_MOVED) != 0).reshape(lead))


def legal_actions_mask_batch(states: np.ndarray) -> np.ndarray:
    """gamemove.py:45 over a batch: bool[N,4] in the order (left, up, right, down)."""
    packed, lead, _ = _boards(states)
    flags = ops.legal_flags(packed).cpu().numpy()
    return (((flags[:, None] >> np.arange(4)) & 1) != 0).reshape(*lead, 4)


def is_done_batch(states: np.ndarray) -> np.ndarray:
    """gameboard.py:328 over a batch (an EMPTY board is not done in the NumPy engine)."""
    a = np.asarray(states)
    packed, lead, _ = _boards(a)
    flags = ops.legal_flags(packed).cpu().numpy()
    return (((flags & FLAG_DONE) != 0) & (a.reshape(-1, 16) != 0).any(axis=1)).reshape(lead)


def after_state_batch(states: np.ndarray) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """gameboard.py:132 over a batch: (children [N,32,4,4], probabilities float64[N,32], counts
    int[N]).  Slot 2k + v of a board = its k-th empty cell (row-major) filled with 2 (v = 0) or 4
    (v = 1), probability {0.9, 0.1} / num_empty; a full board has the single child itself with
    probability 1 (gameboard.py:158-159); unused slots are zero."""
    packed, lead, dtype = _boards(states)
    n = packed.numel()
    children, num_empty, _ = ops.after_state(packed)
    ne = num_empty.cpu().numpy().astype(np.int64)
    counts = np.where(ne > 0, 2 * ne, 1)
    probs = np.zeros((n, 32), dtype=np.float64)
    slot = np.arange(32)[None, :]
    base = np.where(slot % 2 == 0, 0.9, 0.1) / np.maximum(ne, 1)[:, None]
    probs[:] = np.where(slot < counts[:, None], base, 0.0)
    probs[ne == 0, 0] = 1.0
    tiles = ops.unpack(children.reshape(-1)).cpu().numpy().reshape(*lead, 32, 4, 4).astype(dtype, copy=False)
    return tiles, probs.reshape(*lead, 32), counts.reshape(lead)
