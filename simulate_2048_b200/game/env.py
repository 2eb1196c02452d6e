"""B200 drop-in for ``reinforce/game/env.py`` (the batched JAX environment of simulate_2048).

``GameState`` keeps the reference's four fields.  ``board`` is stored PACKED (int64, one word
per env; ``state.tiles`` gives the reference's int32[..., 4, 4] view) so that a batched step is
one kernel over 8 + 4 + 1 + 4 bytes of state per env.  ``batched_step`` is one launch of
``g2048_env_step``, which fuses next_state, is_done, the step counter, the reward accumulator
and the ``info`` dict of env.py:79-133.

PRNG keys: wherever the reference takes ``keys`` (``uint32[N, 2]``), pass

  * ``None`` or an ``int`` seed  -> Philox spawns keyed by (seed, env index, step_count);
  * an ``int64[N]`` tensor       -> one Philox key per env (the analogue of ``jax.random.split``);
  * a ``uint8[N]`` tensor        -> injected spawn bytes (bit-exact replay of the reference).
"""

from __future__ import annotations

from typing import NamedTuple, Optional, Tuple, Union

import torch

from .. import ops
from .._cabi import FLAG_DONE
from . import core

Array = torch.Tensor
Keys = Union[None, int, torch.Tensor]


class GameState(NamedTuple):
    """env.py:33 — board (packed int64), step_count int32, done bool, total_reward float32."""

    board: Array
    step_count: Array
    done: Array
    total_reward: Array

    @property
    def tiles(self) -> Array:
        """The reference's board layout: int32[..., 4, 4] tile values."""
        return core.unpack_boards(self.board)


def _split_keys(keys: Keys, n: int):
    """-> (seed, env_keys, spawns).  The interpretation follows dtype and shape alone:
         None / int                      one Philox key for the batch (seed)
         uint8[n]                        injected spawn bytes
         int64[n]                        one Philox key per env
         (u)int32 / int64 [n, 2]         JAX-style per-env keys (two 32-bit words each)
         anything with 1 or 2 elements   a single key (scalar, or the two words of a JAX key)
    so an int64[2] tensor with n == 2 is two per-env keys, and a uint32[2] tensor is ONE key."""
    if keys is None:
        return 0, None, None
    if isinstance(keys, int):
        return keys, None, None
    if keys.dtype == torch.uint8:
        if keys.numel() != n:
            raise ValueError(f'spawns: expected {n} bytes, got {keys.numel()}')
        return 0, None, keys.reshape(-1)
    if keys.dim() == 2 and keys.shape == (n, 2):
        k = keys.to(torch.int64)
        return 0, ((k[:, 0] & 0xFFFFFFFF) | ((k[:, 1] & 0xFFFFFFFF) << 32)).contiguous(), None
    if keys.dtype == torch.int64 and keys.dim() == 1 and keys.numel() == n:
        return 0, keys.contiguous(), None
    if keys.numel() in (1, 2):
        return core.key_to_seed(keys), None, None
    raise ValueError(f'keys: cannot interpret a {keys.dtype} tensor of shape {tuple(keys.shape)} for {n} envs')


def batched_reset(keys: Keys = None, n: Optional[int] = None, *, spawns: Optional[Array] = None,
                  device=None) -> GameState:
    """env.py:176 — N fresh games.  ``keys`` int64[N] / uint32[N,2] per-env keys, or an int seed
    together with ``n``."""
    if n is None:
        if not isinstance(keys, torch.Tensor):
            raise ValueError('batched_reset: give per-env keys or n')
        n = keys.shape[0]
        device = keys.device if device is None else device
    seed, env_keys, sp = _split_keys(keys, n)
    if spawns is None and sp is not None:
        spawns = sp
    dev = torch.device('cuda' if device is None else device)
    board = ops.reset(n, seed=seed, env_keys=env_keys, spawns=spawns, device=dev)
    return GameState(
        board=board,
        step_count=torch.zeros(n, dtype=torch.int32, device=dev),
        done=torch.zeros(n, dtype=torch.bool, device=dev),
        total_reward=torch.zeros(n, dtype=torch.float32, device=dev),
    )


def batched_step(states: GameState, actions: Array, keys: Keys = None, *, inplace: bool = False
                 ) -> Tuple[GameState, Array, Array, dict]:
    """env.py:211 — one launch for the whole batch.  Returns (states, rewards f32[N],
    dones bool[N], info) with info = {max_tile, step_count, total_reward} (env.py:127-131)."""
    n = states.board.numel()
    if isinstance(keys, int):       # the common call: one seed for the batch
        seed, env_keys, spawns = keys, None, None
    else:
        seed, env_keys, spawns = _split_keys(keys, n)
    if inplace:
        board, sc, tr = states.board, states.step_count, states.total_reward
        done8 = states.done.view(torch.uint8)
    else:
        board, sc, tr = states.board.clone(), states.step_count.clone(), states.total_reward.clone()
        done8 = states.done.to(torch.uint8)
    dev = board.device
    rewards = torch.empty(n, dtype=torch.float32, device=dev)
    max_tile = torch.empty(n, dtype=torch.int32, device=dev)
    ops.env_step(board, sc, done8, tr, actions, spawns, seed=seed, env_keys=env_keys,
                 rewards=rewards, max_tile=max_tile)
    done = done8.view(torch.bool)
    new = GameState(board, sc, done, tr)
    # env.py:104,124: a finished env reports done=True
    return new, rewards, done, {'max_tile': max_tile, 'step_count': sc, 'total_reward': tr}


def batched_get_observation(states: GameState) -> Array:
    """env.py:228 — float32[N, 16].  (A GameState board is packed int64[N]: straight to the kernel,
    without the representation checks and reshapes of the generic core function.)"""
    b = states.board
    if b.dim() == 1 and b.dtype == torch.int64:
        return ops.observe(b, want_max=False)[0]
    return core.encode_observation(b)


def batched_get_legal_actions(states: GameState) -> Array:
    """env.py:245 — bool[N, 4]."""
    b = states.board
    if b.dim() == 1 and b.dtype == torch.int64:
        return ops.legal_actions(b)
    return core.legal_actions_mask(b)


# -- single-env forms (env.py:55, :79, :136, :154): a batch of one -----------------------------

def _one(state: GameState) -> GameState:
    return GameState(state.board.reshape(1), state.step_count.reshape(1), state.done.reshape(1),
                     state.total_reward.reshape(1))


def _scalar(state: GameState) -> GameState:
    return GameState(state.board.reshape(()), state.step_count.reshape(()), state.done.reshape(()),
                     state.total_reward.reshape(()))


def reset(key: Keys = None, *, spawns: Optional[Array] = None, device=None) -> GameState:
    """env.py:55."""
    seed = core.key_to_seed(key)
    return _scalar(batched_reset(seed, 1, spawns=spawns, device=device))


def step(state: GameState, action, key: Keys = None) -> Tuple[GameState, Array, Array, dict]:
    """env.py:79."""
    a = core._actions(action, 1, state.board.device)
    if isinstance(key, torch.Tensor) and key.dtype == torch.uint8:
        k: Keys = key.reshape(1)
    else:
        k = core.key_to_seed(key)
    new, r, d, info = batched_step(_one(state), a, k)
    return _scalar(new), r.reshape(()), d.reshape(()), {k_: v.reshape(()) for k_, v in info.items()}


def get_observation(state: GameState) -> Array:
    """env.py:136 — float32[16]."""
    return core.encode_observation(state.board.reshape(1)).reshape(16)


def get_legal_actions(state: GameState) -> Array:
    """env.py:154 — bool[4]."""
    return core.legal_actions_mask(state.board.reshape(1)).reshape(4)


class Environment:
    """env.py:248 — stateful convenience wrapper.  The reference splits a JAX key per call; here
    the seed stays fixed and the Philox counter advances with the env's own step_count, which
    gives the same property (a fresh, reproducible draw every step)."""

    def __init__(self, seed: int = 0, device=None):
        self.key = seed
        self._episode = 0
        self._device = device
        self.state = self.reset_env()

    def reset_env(self) -> GameState:
        self._episode += 1
        self.state = reset(self.key + 0x9E3779B97F4A7C15 * self._episode, device=self._device)
        return self.state

    def step_env(self, action: int):
        self.state, reward, done, info = step(
            self.state, action, self.key + 0x9E3779B97F4A7C15 * self._episode)
        return self.state, float(reward), bool(done), info

    def observation(self) -> Array:
        return get_observation(self.state)

    def legal_actions(self) -> Array:
        return get_legal_actions(self.state)

    @property
    def board(self) -> Array:
        return self.state.tiles

    @property
    def done(self) -> bool:
        return bool(self.state.done)
