"""Drop-in for ``twentyfortyeight/envs/twentyfortyeight.py`` (class ``TwentyFortyEight``).

This class IS the reference's API for one game: its constructor arguments, properties, ``reset`` /
``step`` signatures and return orders have to be the reference's, so its ten statements necessarily
read like the original's (envs/twentyfortyeight.py:10-140).  What differs is underneath: every
function it calls (``fill_cells``, ``next_state``, ``is_done``, ``encode_flatten``) runs in the CUDA
kernels behind the C ABI.  Each call is a host/device round trip — for throughput use
``VectorTwentyFortyEight`` (vector.py: N lock-step games, one launch per step)."""

from __future__ import annotations

import numpy as np
from numpy import int64, ndarray, zeros

from ... import ops
from ..._cabi import FLAG_DONE, FLAG_MOVED, FLAG_OVERFLOW
from ..core import gameboard as _gb
from ..core._device import pack_host, unpack_host
from ..core.gameboard import fill_cells, is_done, next_state
from ..utils.binary import encode_flatten
from ..utils.normalize import normalize_reward


class TwentyFortyEight:
    """twentyfortyeight.py:10 — 4x4 game with the reference's ``encoded`` (31-way one-hot
    observation, :66-67) and ``normalize`` (log-scaled reward, :80-81) presentation options.
    Other board sizes are out of scope for the packed device format and raise."""

    ACTIONS = {'left': 0, 'up': 1, 'right': 2, 'down': 3}  # twentyfortyeight.py:19

    def __init__(self, size: int = 4, encoded: bool = False, normalize: bool = False):
        if size != 4:
            raise NotImplementedError('the packed device format implements the 4x4 board only')
        self.size = size
        self._encoded = encoded
        self._normalize = normalize
        self._current_state: ndarray = zeros((size, size), dtype=int64)
        self._current_reward: float = 0.0
        self._known = None   # (copy of the board, its flags byte) of the last step: saves is_finished a round trip
        self._io = (np.zeros(1, np.uint64), np.zeros(1, np.uint8),
                    (np.zeros(1, np.uint64), np.zeros(1, np.int32), np.zeros(1, np.uint8)))
        self.reset()

    @property
    def is_finished(self) -> bool:
        k = self._known
        if k is not None and np.array_equal(k[0], self._current_state):
            # gameboard.py:328 — no legal move; the empty board is not "done" in the NumPy engine
            return bool(k[1] & FLAG_DONE) and bool(np.any(k[0]))
        return is_done(self._current_state)

    @property
    def observation(self) -> ndarray:
        if self._encoded:
            return encode_flatten(self._current_state, encodage_size=31)  # twentyfortyeight.py:66-67
        return self._current_state

    @property
    def reward(self) -> float:
        if self._normalize:
            return normalize_reward(self._current_reward)  # twentyfortyeight.py:80-81
        return self._current_reward

    def reset(self, seed: int | None = None) -> ndarray:
        """twentyfortyeight.py:84 — empty board + two tiles, reward 0."""
        self._current_state = zeros((self.size, self.size), dtype=int64)
        self._current_state = fill_cells(state=self._current_state, number_tile=2, seed=seed)
        self._current_reward = 0
        return self.observation

    def step(self, action: int) -> tuple[ndarray, float, bool]:
        """twentyfortyeight.py:106 — (observation, reward of this step, finished).

        One host/device round trip: the board is packed on the host (16 nibbles), ``g2048_step_host``
        runs the transition (module Philox stream, as ``next_state`` without a seed) and returns
        board, reward and the flags byte, from which ``is_finished`` is answered without another
        call.  A board the packed format cannot hold goes through ``next_state`` (which raises)."""
        word = pack_host(self._current_state)
        if word is None or self._current_state.shape != (4, 4):
            self._known = None
            self._current_state, self._current_reward = next_state(state=self._current_state, action=action)
            return self.observation, self.reward, self.is_finished
        boards, actions, out = self._io
        boards[0], actions[0] = word, int(action) % 4
        ops.step_host(boards, actions, None, seed=_gb._GENERATOR.key, step_base=_gb._GENERATOR.next(), out=out)
        flags = int(out[2][0])
        if flags & FLAG_OVERFLOW:
            raise OverflowError('a 32768+32768 merge is not representable in the packed board format')
        if flags & FLAG_MOVED:
            self._current_state = unpack_host(out[0][0], self._current_state.dtype)
            self._current_reward = float(out[1][0])
        else:
            self._current_reward = 0   # gameboard.py:325: (state, 0) — the same array object
        self._known = (self._current_state.copy(), flags)
        return self.observation, self.reward, self.is_finished

    def render(self) -> None:
        for row in self._current_state.tolist():
            print(' \t'.join(map(str, row)))
