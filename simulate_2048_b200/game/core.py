"""B200 drop-in for ``reinforce/game/core.py`` (JAX game core of simulate_2048).

Same function names, argument order and return order.  ``jax.Array`` becomes a CUDA
``torch.Tensor``; boards may be given either as tile values ``int32[..., 4, 4]`` (the reference
layout) or packed ``int64[...]`` (one word per board, see include/g2048.h) and come back in the
form they were given.  A leading batch dimension is accepted everywhere (``jax.vmap`` in the
reference).  A JAX PRNG key is replaced by

  * ``key``: an ``int`` (or scalar / ``(2,)`` uint32 tensor) seeding the counter-based Philox
    stream keyed by (key, env index, step), or
  * ``spawns``: an explicit uint8 tensor of ABI spawn bytes (injected-spawn mode, used for
    bit-exact comparison with the reference, whose threefry stream cannot be reproduced).

Everything runs in the CUDA kernels behind the C ABI; there is no CPU path.
"""

from __future__ import annotations

from typing import Optional, Tuple, Union

import torch

from .. import ops
from .._cabi import FLAG_DONE, FLAG_LEGAL_MASK, FLAG_OVERFLOW

Array = torch.Tensor
Key = Union[int, torch.Tensor, None]

BOARD_SIZE = 4
NUM_ACTIONS = 4

_BITS = None


def _bits(device) -> Array:
    global _BITS
    if _BITS is None or _BITS.device != device:
        _BITS = torch.tensor([1, 2, 4, 8], dtype=torch.uint8, device=device)
    return _BITS


# -- representation helpers ---------------------------------------------------------------------

def is_packed(board: Array) -> bool:
    """True for packed int64 boards, False for [..., 4, 4] tile-value boards."""
    return not (board.dim() >= 2 and tuple(board.shape[-2:]) == (4, 4))


def pack_boards(board: Array) -> Array:
    """int[..., 4, 4] tile values -> packed int64[...] (g2048_pack)."""
    if is_packed(board):
        if board.dtype != torch.int64:
            raise TypeError('packed boards must be int64')
        return board.contiguous()
    lead = board.shape[:-2]
    return ops.pack(board.to(torch.int32).contiguous().reshape(-1, 16)).reshape(lead)


def unpack_boards(board: Array) -> Array:
    """packed int64[...] -> int32[..., 4, 4] tile values (g2048_unpack)."""
    if not is_packed(board):
        return board
    return ops.unpack(board.contiguous().reshape(-1)).reshape(*board.shape, 4, 4)


def _lead(state: Array):
    return state.shape if is_packed(state) else state.shape[:-2]


def _like(packed: Array, template: Array) -> Array:
    packed = packed.reshape(_lead(template))
    return packed if is_packed(template) else unpack_boards(packed).to(template.dtype)


def _actions(action, n: int, device) -> Array:
    if isinstance(action, torch.Tensor) and action.dtype == torch.uint8 and action.is_cuda and action.numel() == n:
        return action.reshape(-1)   # the kernels read only the low two bits (rot90(k) is modulo 4)
    if isinstance(action, int):
        return torch.full((n,), action & 3, dtype=torch.uint8, device=device)
    a = torch.as_tensor(action, device=device).reshape(-1)
    if a.numel() == 1 and n != 1:
        a = a.expand(n)
    return (a & 3).to(torch.uint8).contiguous()


def key_to_seed(key: Key) -> int:
    """int / scalar tensor / JAX-style (2,) uint32 key -> 64-bit Philox key."""
    if key is None:
        return 0
    if isinstance(key, int):
        return key & (2**64 - 1)
    k = torch.as_tensor(key).reshape(-1)
    if k.numel() == 1:
        return int(k.item()) & (2**64 - 1)
    if k.numel() == 2:
        return (int(k[0].item()) & 0xFFFFFFFF) | ((int(k[1].item()) & 0xFFFFFFFF) << 32)
    raise ValueError('key: expected an int, a scalar tensor or a (2,) uint32 key')


def _raise_on_overflow(flags: Array) -> None:
    if bool((flags & FLAG_OVERFLOW).any().item()):
        raise OverflowError('a 32768+32768 merge is not representable in the packed board format')


# -- core.py:43 merge_row / :104 slide_and_merge / :129 latent_state ---------------------------

def latent_state(state: Array, action) -> Tuple[Array, Array]:
    """core.py:129 — afterstate (no spawn) and int32 reward of ``action`` (0 L, 1 U, 2 R, 3 D)."""
    packed = pack_boards(state).reshape(-1)
    n = packed.numel()
    out = ops.expand(packed, children='none')
    _raise_on_overflow(out.mask)
    idx = _actions(action, n, packed.device).to(torch.int64).unsqueeze(1)
    after = out.after.gather(1, idx).squeeze(1)
    reward = out.rewards.gather(1, idx).squeeze(1)
    return _like(after, state), reward.reshape(_lead(state))


def slide_and_merge(board: Array) -> Tuple[Array, Array]:
    """core.py:104 — slide the whole board left.  Returns (board, score): board FIRST."""
    return latent_state(board, 0)


def merge_row(row: Array) -> Tuple[Array, Array]:
    """core.py:43 — merge one row ([4]) or a batch of rows ([..., 4]) to the left."""
    rows = row.reshape(-1, 4).to(torch.int32)
    boards = torch.zeros((rows.shape[0], 4, 4), dtype=torch.int32, device=row.device)
    boards[:, 0, :] = rows
    after, score = latent_state(boards, 0)
    return after[:, 0, :].reshape(row.shape).to(row.dtype), score.reshape(row.shape[:-1])


# -- core.py:173 fill_cells / :234 next_state / :324 create_initial_board ----------------------

def _need_randomness(name: str, key: Key, spawns) -> None:
    """The reference requires a PRNG key (core.py:173, :234); a silent default would give every
    call the same Philox counter (seed 0, env i, step 0), i.e. the same draw for a whole game."""
    if key is None and spawns is None:
        raise TypeError(f'{name}: pass key= (an int or a (2,) uint32 key) or spawns= (injected spawn bytes)')


def fill_cells(state: Array, key: Key = None, *, spawns: Optional[Array] = None, step: int = 0,
               tag: int = 0) -> Array:
    """core.py:173 — add one tile (2 w.p. 0.9, 4 w.p. 0.1) on a uniformly chosen empty cell; a full
    board is returned unchanged.  ``key`` (or ``spawns``) is mandatory, as in the reference; split
    or advance it between calls (``step=`` advances the counter under one key)."""
    _need_randomness('fill_cells', key, spawns)
    packed = pack_boards(state).reshape(-1)
    new = ops.spawn(packed, spawns, seed=key_to_seed(key), step=step, tag=tag)
    return _like(new, state)


def next_state(state: Array, action, key: Key = None, *, spawns: Optional[Array] = None,
               step: int = 0) -> Tuple[Array, Array]:
    """core.py:234 — full transition: afterstate, then one new tile iff the board changed; the
    reward is float32 and 0 for an illegal move (core.py:270).  ``key`` (or ``spawns``) is
    mandatory, as in the reference."""
    _need_randomness('next_state', key, spawns)
    packed = pack_boards(state).reshape(-1)
    out = ops.step(packed, _actions(action, packed.numel(), packed.device), spawns,
                   seed=key_to_seed(key), step_base=step)
    _raise_on_overflow(out.flags)
    return _like(out.boards, state), out.rewards.to(torch.float32).reshape(_lead(state))


def create_initial_board(key: Key = None, *, n: Optional[int] = None,
                         spawns: Optional[Array] = None, device=None) -> Array:
    """core.py:324 — zeros + two sequential fill_cells.  Returns int32[4,4] (or [n,4,4])."""
    boards = ops.reset(1 if n is None else n, seed=key_to_seed(key), spawns=spawns, device=device)
    tiles = unpack_boards(boards)
    return tiles[0] if n is None else tiles


# -- core.py:275 legal_actions_mask / :305 is_done ----------------------------------------------

def legal_actions_mask(state: Array) -> Array:
    """core.py:275 — bool[..., 4] in the order [left, up, right, down] (one launch: the kernel
    writes the bool array itself)."""
    packed = pack_boards(state).reshape(-1)
    return ops.legal_actions(packed).reshape(*_lead(state), 4)


def is_done(state: Array) -> Array:
    """core.py:305 — True when no action is legal (an empty board counts as done, as in JAX)."""
    packed = pack_boards(state).reshape(-1)
    flags = ops.legal_flags(packed)
    return ((flags & FLAG_DONE) != 0).reshape(_lead(state))


# -- core.py:346 encode_observation / :427 max_tile / :445 count_empty -------------------------

def encode_observation(board: Array) -> Array:
    """core.py:346 — float32[..., 16] = log2(tile)/16, 0 for an empty cell."""
    packed = pack_boards(board).reshape(-1)
    obs, _ = ops.observe(packed, want_max=False)
    return obs.reshape(*_lead(board), 16)


# -- core.py:381 sample_action ------------------------------------------------------------------

def _policy_batch(policy: Array, legal_mask: Array):
    pol = policy.to(torch.float32).reshape(-1, NUM_ACTIONS).contiguous()
    legal = legal_mask.to(torch.bool).reshape(-1, NUM_ACTIONS).contiguous()
    if pol.shape != legal.shape:
        raise ValueError('policy and legal_mask must have the same shape [..., 4]')
    return pol, legal


def action_distribution(temperature: float, policy: Array, legal_mask: Array) -> Array:
    """The distribution core.py:401-421 samples from: illegal actions masked, renormalised (uniform
    over the legal actions when the mass is below 1e-8), then softmax(log(p + 1e-8) / temperature),
    or a one-hot of the argmax when temperature < 0.01.  float32[..., 4] (g2048_sample_action)."""
    pol, legal = _policy_batch(policy, legal_mask)
    _, probs = ops.sample_action(pol, legal, temperature, want_probs=True)
    return probs.reshape(policy.shape)


def sample_action(key: Key, temperature: float, policy: Array, legal_mask: Array) -> Array:
    """core.py:381 — an action drawn from ``action_distribution`` (int64[...]; a scalar tensor for a
    single policy), one kernel launch for the whole batch.  The JAX key becomes the Philox key of the
    draw (word of (key, row index, 0, tag 4)): the threefry draw itself cannot be reproduced, the
    distribution is the reference's."""
    pol, legal = _policy_batch(policy, legal_mask)
    picks = ops.sample_action(pol, legal, temperature, seed=key_to_seed(key))
    return picks.to(torch.int64).reshape(policy.shape[:-1])


def max_tile(board: Array) -> Array:
    """core.py:427 — the largest tile VALUE (int32), 0 for an empty board."""
    packed = pack_boards(board).reshape(-1)
    _, mx = ops.observe(packed, want_max=True)
    mx = mx.to(torch.int32)
    return torch.where(mx > 0, torch.ones_like(mx) << mx, torch.zeros_like(mx)).reshape(_lead(board))


def count_empty(board: Array) -> Array:
    """core.py:445 — number of empty cells."""
    packed = pack_boards(board).reshape(-1)
    obs, _ = ops.observe(packed, want_max=False)
    return (obs == 0).sum(dim=1).to(torch.int32).reshape(_lead(board))


__all__ = [
    'BOARD_SIZE', 'NUM_ACTIONS', 'merge_row', 'slide_and_merge', 'latent_state', 'fill_cells',
    'next_state', 'legal_actions_mask', 'is_done', 'create_initial_board', 'encode_observation',
    'max_tile', 'count_empty', 'sample_action', 'action_distribution', 'pack_boards', 'unpack_boards', 'is_packed', 'key_to_seed',
    'FLAG_LEGAL_MASK',
]
