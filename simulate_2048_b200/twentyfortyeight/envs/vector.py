"""``VectorTwentyFortyEight`` — N lock-step ``TwentyFortyEight`` games behind the reference's
env surface (``envs/twentyfortyeight.py:10-140``: ``reset``, ``step``, ``observation``, ``reward``,
``is_finished``), one kernel launch per step for the whole batch.

The single-game mirror class pays a host/device round trip per call; this class is the form in
which the drop-in is fast for the reference's own use (BASELINE configs[0]: 1024 random-policy
games): state lives on the device as packed boards, a step copies N action bytes in and N boards,
rewards and flag bytes out through pinned buffers.
"""

from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch

from ... import ops
from ..._cabi import FLAG_DONE, FLAG_LEGAL_MASK
from ..core._device import device


class VectorTwentyFortyEight:
    ACTIONS = {'left': 0, 'up': 1, 'right': 2, 'down': 3}  # twentyfortyeight.py:19

    def __init__(self, num_envs: int, encoded: bool = False, normalize: bool = False, seed: Optional[int] = None):
        self.num_envs = n = int(num_envs)
        self.size = 4
        self._encoded, self._normalize = encoded, normalize
        dev = device()
        self._board = torch.zeros(n, dtype=torch.int64, device=dev)
        self._steps = torch.zeros(n, dtype=torch.int32, device=dev)
        self._done = torch.zeros(n, dtype=torch.uint8, device=dev)
        self._total = torch.zeros(n, dtype=torch.float32, device=dev)
        self._act = torch.zeros(n, dtype=torch.uint8, device=dev)
        self._rew = torch.zeros(n, dtype=torch.float32, device=dev)
        self._flags = torch.zeros(n, dtype=torch.uint8, device=dev)
        self._tiles = torch.zeros((n, 4, 4), dtype=torch.int32, device=dev)
        # pinned host mirrors (g2048_host_alloc), filled by asynchronous copies
        self._h_act = ops.pinned_empty(n, torch.uint8)
        self._h_tiles = ops.pinned_empty(n * 16, torch.int32)
        self._h_rew = ops.pinned_empty(n, torch.float32)
        self._h_flags = ops.pinned_empty(n, torch.uint8)
        self._t_act = torch.from_numpy(self._h_act)
        self._t_tiles = torch.from_numpy(self._h_tiles).reshape(n, 4, 4)
        self._t_rew = torch.from_numpy(self._h_rew)
        self._t_flags = torch.from_numpy(self._h_flags)
        self._episode = 0
        self._seed = int.from_bytes(np.random.bytes(8), 'little') if seed is None else int(seed)
        self.reset(seed)

    # -- reference surface -------------------------------------------------------------------
    def reset(self, seed: Optional[int] = None) -> np.ndarray:
        """twentyfortyeight.py:84 for every game: empty board + two tiles, reward 0.  Game i draws
        from the Philox stream keyed by (seed, i)."""
        if seed is not None:
            self._seed = int(seed)
        self._episode += 1
        self._key = (self._seed + 0x9E3779B97F4A7C15 * self._episode) & (2**64 - 1)
        self._board.copy_(ops.reset(self.num_envs, seed=self._key, device=self._board.device))
        self._steps.zero_()
        self._done.zero_()
        self._total.zero_()
        self._rew.zero_()
        self._flags.copy_(ops.legal_flags(self._board))
        self._pull()
        return self.observation

    def step(self, actions) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """twentyfortyeight.py:106 for every game: (observations, rewards of this step, finished).
        A finished game is left unchanged with reward 0 (its ``next_state`` has no legal move)."""
        self._h_act[:] = np.asarray(actions, dtype=np.uint8).reshape(-1) & 3
        self._act.copy_(self._t_act, non_blocking=True)
        ops.env_step(self._board, self._steps, self._done, self._total, self._act, None, seed=self._key,
                     rewards=self._rew, flags=self._flags)
        self._pull()
        return self.observation, self.reward, self.is_finished

    @property
    def observation(self) -> np.ndarray:
        """int64[N,4,4] tile values (or the 31-way one-hot rows, twentyfortyeight.py:66-67)."""
        if self._encoded:
            return ops.observe_onehot(self._board, 31).reshape(self.num_envs, 16 * 31).cpu().numpy()
        return self._h_tiles.reshape(self.num_envs, 4, 4).astype(np.int64)

    @property
    def reward(self) -> np.ndarray:
        r = self._h_rew.astype(np.float64)
        if self._normalize:  # utils/normalize.py:6
            return np.where(r > 0, np.log2(np.maximum(r, 1.0)) / 16.0, 0.0)
        return r

    @property
    def is_finished(self) -> np.ndarray:
        return (self._h_flags & FLAG_DONE) != 0

    # -- batch extras --------------------------------------------------------------------------
    @property
    def legal_actions_mask(self) -> np.ndarray:
        """bool[N,4] (gamemove.py:45) of the current boards — from the flag byte the step returned."""
        return ((self._h_flags[:, None] >> np.arange(4)) & 1) != 0

    @property
    def total_reward(self) -> np.ndarray:
        return self._total.cpu().numpy()

    def _pull(self) -> None:
        ops.unpack_into(self._board, self._tiles)
        self._t_tiles.copy_(self._tiles, non_blocking=True)
        self._t_rew.copy_(self._rew, non_blocking=True)
        self._t_flags.copy_(self._flags, non_blocking=True)
        torch.cuda.current_stream().synchronize()
