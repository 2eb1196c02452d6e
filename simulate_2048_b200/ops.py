"""Tensor-level operators over the C ABI (include/g2048.h).

Every function takes CUDA ``torch`` tensors (torch is only the owner of device memory and
streams here), checks dtype / device / contiguity, and launches the matching ``g2048_*`` entry
point on the current CUDA stream.  Boards are packed ``int64`` tensors (bit pattern of the
``uint64`` word: nibble ``i`` = log2 of flat cell ``i = 4r + c``).  Nothing here computes on
the CPU and nothing falls back: a missing library or a non-CUDA tensor raises.
"""

from __future__ import annotations

from typing import NamedTuple, Optional

import numpy as np
import torch

from . import _cabi
from ._cabi import check, lib

__all__ = [
    'StepOut', 'ExpandOut', 'RolloutOut', 'step', 'env_step', 'selfplay_step', 'sample_action', 'legal_flags', 'legal_actions', 'expand', 'reset',
    'observe', 'observe_onehot', 'nstep_returns', 'expect', 'sample_children', 'expect_sampled', 'spawn', 'after_state', 'rollout', 'pack', 'unpack', 'unpack_into', 'step_host', 'row_table', 'set_impl', 'set_tuning', 'init',
    'pinned_empty', 'host_numa_info', 'step_host_copies', 'shutdown',
]


try:  # raw handle of the current stream without building a torch.cuda.Stream object (~1 us less per call)
    _raw_stream = torch._C._cuda_getCurrentRawStream
except AttributeError:  # pragma: no cover - older / newer torch without the private helper
    _raw_stream = None


def _stream() -> int:
    if _raw_stream is not None:
        return _raw_stream(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream


class _NoSwitch:
    def __enter__(self):
        return None

    def __exit__(self, *a):
        return False


_NO_SWITCH = _NoSwitch()


def _on(device):
    """Device guard that costs nothing when ``device`` is already current (``torch.cuda.device`` is
    ~5 us per call, as much as a small launch)."""
    if device.index is None or device.index == torch.cuda.current_device():
        return _NO_SWITCH
    return torch.cuda.device(device)


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _req(t: torch.Tensor, dtype: torch.dtype, name: str, n: Optional[int] = None) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f'{name}: expected a torch.Tensor, got {type(t).__name__}')
    if not t.is_cuda:
        raise ValueError(f'{name}: must be a CUDA tensor (simulate_2048_b200 has no CPU path)')
    if t.dtype != dtype:
        raise TypeError(f'{name}: expected dtype {dtype}, got {t.dtype}')
    if not t.is_contiguous():
        raise ValueError(f'{name}: must be contiguous')
    if n is not None and t.numel() != n:
        raise ValueError(f'{name}: expected {n} elements, got {t.numel()}')
    return t


def _opt(t, dtype, name, n=None):
    return None if t is None else _req(t, dtype, name, n)


def _u8(actions: torch.Tensor, name: str, n: int) -> torch.Tensor:
    """Accept uint8 directly; cast int32/int64 actions (the reference passes int32)."""
    if actions.dtype != torch.uint8:
        actions = actions.to(torch.uint8)
    return _req(actions.contiguous(), torch.uint8, name, n)


def init(device: Optional[int] = None) -> None:
    """Build the per-device row table now (g2048_init)."""
    check(lib().g2048_init(-1 if device is None else int(device)))


def set_impl(impl: int) -> None:
    check(lib().g2048_set_impl(int(impl)))


def set_tuning(step_threads: int, envs_per_thread: int = 4) -> None:
    check(lib().g2048_set_tuning(int(step_threads), int(envs_per_thread)))


class StepOut(NamedTuple):
    boards: torch.Tensor   # int64[N] packed
    rewards: torch.Tensor  # int32[N]
    flags: torch.Tensor    # uint8[N]  (see include/g2048.h G2048_FLAG_*)


def step(boards: torch.Tensor, actions: torch.Tensor, spawns: Optional[torch.Tensor] = None, *,
         seed: int = 0, env_offset: int = 0, step_ctr: Optional[torch.Tensor] = None,
         step_base: int = 0, env_keys: Optional[torch.Tensor] = None,
         done_in: Optional[torch.Tensor] = None, out: Optional[StepOut] = None) -> StepOut:
    """Raw batched transition (g2048_step): gameboard.py:291 / core.py:234 next_state fused with
    the legal mask and done flag of the new board.  ``spawns=None`` selects Philox spawns."""
    n = boards.numel()
    _req(boards, torch.int64, 'boards')
    actions = _u8(actions, 'actions', n)
    spawns = None if spawns is None else _u8(spawns, 'spawns', n)
    _opt(step_ctr, torch.int32, 'step_ctr', n)
    _opt(env_keys, torch.int64, 'env_keys', n)
    done_in = None if done_in is None else _u8(done_in, 'done_in', n)
    if out is None:
        out = StepOut(torch.empty_like(boards), torch.empty(n, dtype=torch.int32, device=boards.device),
                      torch.empty(n, dtype=torch.uint8, device=boards.device))
    else:
        _req(out.boards, torch.int64, 'out.boards', n)
        _req(out.rewards, torch.int32, 'out.rewards', n)
        _req(out.flags, torch.uint8, 'out.flags', n)
    with _on(boards.device):
        check(lib().g2048_step(_ptr(boards), _ptr(actions), _ptr(spawns), seed & (2**64 - 1),
                               env_offset, _ptr(step_ctr), step_base & 0xFFFFFFFF, _ptr(env_keys),
                               _ptr(done_in), _ptr(out.boards), _ptr(out.rewards), _ptr(out.flags),
                               n, _stream()))
    return out


def env_step(boards: torch.Tensor, step_count: torch.Tensor, done: torch.Tensor,
             total_reward: torch.Tensor, actions: torch.Tensor,
             spawns: Optional[torch.Tensor] = None, *, seed: int = 0, env_offset: int = 0,
             env_keys: Optional[torch.Tensor] = None, rewards: Optional[torch.Tensor] = None,
             flags: Optional[torch.Tensor] = None, obs: Optional[torch.Tensor] = None,
             max_tile: Optional[torch.Tensor] = None, legal: Optional[torch.Tensor] = None) -> None:
    """In-place GameState step (g2048_env_step; env.py:79-133).  Optional outputs describe the
    NEW state: rewards f32[N], flags u8[N], obs f32[N,16], max_tile i32[N], legal bool[N,4]."""
    n = boards.numel()
    _req(boards, torch.int64, 'boards')
    _req(step_count, torch.int32, 'step_count', n)
    _req(done, torch.uint8, 'done', n)
    _req(total_reward, torch.float32, 'total_reward', n)
    actions = _u8(actions, 'actions', n)
    spawns = None if spawns is None else _u8(spawns, 'spawns', n)
    _opt(env_keys, torch.int64, 'env_keys', n)
    _opt(rewards, torch.float32, 'rewards', n)
    _opt(flags, torch.uint8, 'flags', n)
    _opt(obs, torch.float32, 'obs', n * 16)
    _opt(max_tile, torch.int32, 'max_tile', n)
    _opt(legal, torch.bool, 'legal', n * 4)
    with _on(boards.device):
        check(lib().g2048_env_step(_ptr(boards), _ptr(step_count), _ptr(done), _ptr(total_reward),
                                   _ptr(actions), _ptr(spawns), seed & (2**64 - 1), env_offset,
                                   _ptr(env_keys), _ptr(rewards), _ptr(flags), _ptr(obs),
                                   _ptr(max_tile), _ptr(legal), n, _stream()))


def sample_action(policy: torch.Tensor, legal: torch.Tensor, temperature: float = 1.0, *, seed: int = 0,
                  env_offset: int = 0, step_ctr: Optional[torch.Tensor] = None, step_base: int = 0,
                  want_probs: bool = False):
    """core.py:381 sample_action for a batch (g2048_sample_action): uint8[N] actions drawn from
    ``policy`` f32[N,4] under ``legal`` bool[N,4]; with ``want_probs`` also the f32[N,4] distribution."""
    n = policy.numel() // 4
    _req(policy, torch.float32, 'policy', n * 4)
    _req(legal, torch.bool, 'legal', n * 4)
    _opt(step_ctr, torch.int32, 'step_ctr', n)
    actions = torch.empty(n, dtype=torch.uint8, device=policy.device)
    probs = torch.empty((n, 4), dtype=torch.float32, device=policy.device) if want_probs else None
    with _on(policy.device):
        check(lib().g2048_sample_action(_ptr(policy), _ptr(legal), float(temperature), seed & (2**64 - 1),
                                        env_offset, _ptr(step_ctr), step_base & 0xFFFFFFFF, _ptr(actions),
                                        _ptr(probs), n, _stream()))
    return (actions, probs) if want_probs else actions


def selfplay_step(boards: torch.Tensor, step_count: torch.Tensor, done: torch.Tensor,
                  total_reward: torch.Tensor, policy: Optional[torch.Tensor] = None, *,
                  values: Optional[torch.Tensor] = None, actions: Optional[torch.Tensor] = None,
                  spawns: Optional[torch.Tensor] = None, temperature: float = 1.0, seed: int = 0,
                  env_offset: int = 0, record=None, rewards=None, norm_rewards=None, obs=None, legal=None,
                  max_tile=None, actions_out=None, probs_out=None) -> None:
    """One self-play iteration in one call (g2048_selfplay_step): sample an action per env from
    ``policy`` f32[N,4] (core.py:381 sample_action semantics; ``actions`` u8[N] overrides the draw),
    append the step to ``record`` (a ``selfplay.TrajectoryBuffer``: row = its device-side counter),
    apply it in place as ``env_step`` does, and write the optional outputs for the NEW state."""
    n = boards.numel()
    _req(boards, torch.int64, 'boards')
    _req(step_count, torch.int32, 'step_count', n)
    _req(done, torch.uint8, 'done', n)
    _req(total_reward, torch.float32, 'total_reward', n)
    _opt(policy, torch.float32, 'policy', n * 4)
    _opt(values, torch.float32, 'values', n)
    actions = None if actions is None else _u8(actions, 'actions', n)
    spawns = None if spawns is None else _u8(spawns, 'spawns', n)
    if policy is None and actions is None:
        raise ValueError('selfplay_step: give a policy to sample from or explicit actions')
    _opt(rewards, torch.float32, 'rewards', n)
    _opt(norm_rewards, torch.float32, 'norm_rewards', n)
    _opt(obs, torch.float32, 'obs', n * 16)
    _opt(legal, torch.bool, 'legal', n * 4)
    _opt(max_tile, torch.int32, 'max_tile', n)
    _opt(actions_out, torch.uint8, 'actions_out', n)
    _opt(probs_out, torch.float32, 'probs_out', n * 4)
    r = record
    if r is not None and r.n != n:
        raise ValueError('record: trajectory buffer is for a different number of envs')
    tr = (lambda name: None) if r is None else (lambda name: _ptr(getattr(r, name)))
    with _on(boards.device):
        check(lib().g2048_selfplay_step(
            _ptr(boards), _ptr(step_count), _ptr(done), _ptr(total_reward), _ptr(policy), _ptr(values),
            _ptr(actions), _ptr(spawns), float(temperature), seed & (2**64 - 1), env_offset,
            tr('t_dev'), 0 if r is None else r.capacity, tr('obs'), tr('legal'), tr('actions'), tr('rewards'),
            tr('policies'), tr('values'), tr('active'), tr('lengths'), _ptr(rewards), _ptr(norm_rewards),
            _ptr(obs), _ptr(legal), _ptr(max_tile), _ptr(actions_out), _ptr(probs_out), n, _stream()))


def legal_actions(boards: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """bool[N,4] legal-action mask straight from the kernel (g2048_legal_actions; core.py:275)."""
    _req(boards, torch.int64, 'boards')
    n = boards.numel()
    if out is None:
        out = torch.empty((n, 4), dtype=torch.bool, device=boards.device)
    else:
        _req(out, torch.bool, 'out', n * 4)
    with _on(boards.device):
        check(lib().g2048_legal_actions(_ptr(boards), _ptr(out), None, n, _stream()))
    return out


def legal_flags(boards: torch.Tensor) -> torch.Tensor:
    """uint8[N]: bits 0-3 legal (L,U,R,D), bit 4 done (g2048_mask; gamemove.py:45, core.py:275)."""
    _req(boards, torch.int64, 'boards')
    flags = torch.empty(boards.numel(), dtype=torch.uint8, device=boards.device)
    with _on(boards.device):
        check(lib().g2048_mask(_ptr(boards), _ptr(flags), boards.numel(), _stream()))
    return flags


class ExpandOut(NamedTuple):
    after: torch.Tensor        # int64[N,4] afterstates (gameboard.py:105)
    rewards: torch.Tensor      # int32[N,4]
    mask: torch.Tensor         # uint8[N] legal | done | overflow
    num_empty: torch.Tensor    # uint8[N,4]
    child_count: torch.Tensor  # uint8[N,4]
    children: Optional[torch.Tensor]       # int64[N,4,16,2] dense or int64[total or capacity] compact
    child_offsets: Optional[torch.Tensor]  # int64[N] (compact)
    total: Optional[torch.Tensor] = None   # int64[2] on the device (compact): {children, overflowed}


CHILDREN_PER_BOARD_MAX = 120  # 4 actions x 15 empty cells x 2 values


def _expand_compact(boards, after, rewards, mask, num_empty, child_count, offsets, children, total):
    check(lib().g2048_expand_compact(_ptr(boards), _ptr(after), _ptr(rewards), _ptr(mask), _ptr(num_empty),
                                     _ptr(child_count), _ptr(offsets), _ptr(children), children.numel(),
                                     _ptr(total), boards.numel(), _stream()))


def expand(boards: torch.Tensor, children: str = 'none', zero_fill: bool = True,
           out: Optional[ExpandOut] = None, capacity: Optional[int] = None,
           child_offsets: Optional[torch.Tensor] = None) -> ExpandOut:
    """MCTS leaf expansion (g2048_expand / g2048_expand_compact): mask, the four afterstates +
    rewards and every chance child of each legal afterstate (gameboard.py:132 after_state order:
    cell-major, 2 then 4; probability of a child = {0.9, 0.1} / num_empty).

    ``children='dense'`` writes every slot (those beyond ``child_count`` hold 0: no memset needed);
    ``zero_fill`` is kept for callers of the earlier signature and ignored.

    ``children='compact'`` packs the children of board i at ``child_offsets[i]``.  The offsets are
    computed on the device in the same call (no host round trip):
      * default: one launch into a buffer sized from an estimate, then ONE host read of the total
        to return an exactly sized ``children`` tensor (a second launch only if the estimate was
        too small);
      * ``capacity=k`` or ``out=`` (buffers of an earlier call, any board contents): fully
        asynchronous; ``children`` keeps its capacity, ``total`` (device int64[2]) holds the number
        of valid children and an overflow flag — a board that does not fit writes nothing;
      * ``child_offsets=`` (caller-supplied, with ``out=``): the offsets are an INPUT and the
        children buffer must be large enough for them (g2048_expand)."""
    _req(boards, torch.int64, 'boards')
    n = boards.numel()
    dev = boards.device
    layout = {'none': _cabi.CHILDREN_NONE, 'dense': _cabi.CHILDREN_DENSE,
              'compact': _cabi.CHILDREN_COMPACT}[children]
    if out is not None:  # reuse a previous result's buffers: no allocation
        if out.after.shape[0] != n or (layout != _cabi.CHILDREN_NONE and out.children is None):
            raise ValueError('out: buffers do not match this call')
        with _on(dev):
            if layout == _cabi.CHILDREN_COMPACT and child_offsets is None:
                total = out.total if out.total is not None else torch.empty(2, dtype=torch.int64, device=dev)
                _req(out.child_offsets, torch.int64, 'out.child_offsets', n)
                _expand_compact(boards, out.after, out.rewards, out.mask, out.num_empty, out.child_count,
                                out.child_offsets, out.children, total)
                return out._replace(total=total)
            if layout == _cabi.CHILDREN_COMPACT:
                _req(child_offsets, torch.int64, 'child_offsets', n)
                out = out._replace(child_offsets=child_offsets)
            check(lib().g2048_expand(_ptr(boards), _ptr(out.after), _ptr(out.rewards), _ptr(out.mask),
                                     _ptr(out.num_empty), _ptr(out.child_count), _ptr(out.children),
                                     _ptr(out.child_offsets), layout, n, _stream()))
        return out
    after = torch.empty((n, 4), dtype=torch.int64, device=dev)
    rewards = torch.empty((n, 4), dtype=torch.int32, device=dev)
    mask = torch.empty(n, dtype=torch.uint8, device=dev)
    num_empty = torch.empty((n, 4), dtype=torch.uint8, device=dev)
    child_count = torch.empty((n, 4), dtype=torch.uint8, device=dev)
    ch = off = None
    with _on(dev):
        if layout == _cabi.CHILDREN_COMPACT:
            off = torch.empty(n, dtype=torch.int64, device=dev)
            total = torch.zeros(2, dtype=torch.int64, device=dev)
            if n == 0:
                return ExpandOut(after, rewards, mask, num_empty, child_count,
                                 torch.empty(0, dtype=torch.int64, device=dev), off, total)
            fixed = capacity is not None
            cap = int(capacity) if fixed else min(CHILDREN_PER_BOARD_MAX * n, 72 * n + 4096)
            ch = torch.empty(max(cap, 1), dtype=torch.int64, device=dev)
            _expand_compact(boards, after, rewards, mask, num_empty, child_count, off, ch, total)
            if fixed:
                return ExpandOut(after, rewards, mask, num_empty, child_count, ch, off, total)
            count, over = (int(x) for x in total.tolist())  # the one host read: sizes the result
            if over:  # the estimate was too small: once more with the exact size
                ch = torch.empty(count, dtype=torch.int64, device=dev)
                _expand_compact(boards, after, rewards, mask, num_empty, child_count, off, ch, total)
            return ExpandOut(after, rewards, mask, num_empty, child_count, ch[:count], off, total)
        if layout == _cabi.CHILDREN_DENSE:
            ch = torch.empty((n, 4, 16, 2), dtype=torch.int64, device=dev)
        check(lib().g2048_expand(_ptr(boards), _ptr(after), _ptr(rewards), _ptr(mask),
                                 _ptr(num_empty), _ptr(child_count), _ptr(ch), _ptr(off), layout, n,
                                 _stream()))
    return ExpandOut(after, rewards, mask, num_empty, child_count, ch, off)


def after_state(boards: torch.Tensor, want_children: bool = True):
    """Chance children of the given boards (g2048_after_state; gameboard.py:132/:174).
    Returns (children int64[N,16,2] or None, num_empty uint8[N], empty_cells int16[N])."""
    _req(boards, torch.int64, 'boards')
    n = boards.numel()
    dev = boards.device
    ch = torch.empty((n, 16, 2), dtype=torch.int64, device=dev) if want_children else None  # every slot is written
    ne = torch.empty(n, dtype=torch.uint8, device=dev)
    ec = torch.empty(n, dtype=torch.int16, device=dev)
    with _on(dev):
        check(lib().g2048_after_state(_ptr(boards), _ptr(ch), _ptr(ne), _ptr(ec), n, _stream()))
    return ch, ne, ec


def reset(n: int, *, seed: int = 0, env_offset: int = 0, env_keys: Optional[torch.Tensor] = None,
          spawns: Optional[torch.Tensor] = None, device=None) -> torch.Tensor:
    """New boards with two tiles (g2048_reset; core.py:324, twentyfortyeight.py:84)."""
    dev = torch.device('cuda' if device is None else device)
    _opt(env_keys, torch.int64, 'env_keys', n)
    if spawns is not None:
        spawns = _u8(spawns, 'spawns', 2 * n)
    boards = torch.empty(n, dtype=torch.int64, device=dev)
    with _on(dev):
        check(lib().g2048_reset(seed & (2**64 - 1), env_offset, _ptr(env_keys), _ptr(spawns),
                                _ptr(boards), n, _stream()))
    return boards


def spawn(boards: torch.Tensor, spawns: Optional[torch.Tensor] = None, *, seed: int = 0,
          env_offset: int = 0, env_keys: Optional[torch.Tensor] = None, step: int = 0,
          tag: int = 0, flags: Optional[torch.Tensor] = None) -> torch.Tensor:
    """One new tile per board (g2048_spawn; gameboard.py:247 / core.py:173 fill_cells).  Returns
    a new tensor; a full board comes back unchanged."""
    _req(boards, torch.int64, 'boards')
    n = boards.numel()
    spawns = None if spawns is None else _u8(spawns, 'spawns', n)
    _opt(env_keys, torch.int64, 'env_keys', n)
    _opt(flags, torch.uint8, 'flags', n)
    out = boards.clone()
    with _on(boards.device):
        check(lib().g2048_spawn(_ptr(out), _ptr(spawns), seed & (2**64 - 1), env_offset,
                                _ptr(env_keys), step & 0xFFFFFFFF, tag & 0xFFFFFFFF, _ptr(flags), n,
                                _stream()))
    return out


def observe(boards: torch.Tensor, want_max: bool = True):
    """f32[N,16] observation = exponent/16 (core.py:346) and uint8[N] max exponent."""
    _req(boards, torch.int64, 'boards')
    n = boards.numel()
    obs = torch.empty((n, 16), dtype=torch.float32, device=boards.device)
    mx = torch.empty(n, dtype=torch.uint8, device=boards.device) if want_max else None
    with _on(boards.device):
        check(lib().g2048_observe(_ptr(boards), _ptr(obs), _ptr(mx), n, _stream()))
    return obs, mx


def observe_onehot(boards: torch.Tensor, width: int = 31) -> torch.Tensor:
    """uint8[N,16,width] one-hot of log2(tile) per cell (g2048_observe_onehot; binary.py:52
    encode_flatten as used by TwentyFortyEight(encoded=True), width 31)."""
    _req(boards, torch.int64, 'boards')
    n = boards.numel()
    out = torch.empty((n, 16, width), dtype=torch.uint8, device=boards.device)
    with _on(boards.device):
        check(lib().g2048_observe_onehot(_ptr(boards), _ptr(out), int(width), n, _stream()))
    return out


def expect(child_values: torch.Tensor, ex: 'ExpandOut', discount: float = 1.0,
           illegal_value: float = float('-inf'), use_rewards: bool = True) -> torch.Tensor:
    """float32[N,4] action values of a depth-1 expectimax backup (g2048_expect): reward + discount *
    E[value of the chance child], with the after_state probabilities of gameboard.py:132-171.
    ``ex`` is a compact ``expand`` result, ``child_values`` float32[total] one value per child."""
    if ex.children is None or ex.child_offsets is None:
        raise ValueError('expect needs the compact children layout of expand()')
    n = ex.after.shape[0]
    _opt(child_values, torch.float32, 'child_values', ex.children.numel())
    q = torch.empty((n, 4), dtype=torch.float32, device=ex.after.device)
    with _on(ex.after.device):
        check(lib().g2048_expect(_ptr(child_values), _ptr(ex.child_count), _ptr(ex.child_offsets),
                                 _ptr(ex.rewards) if use_rewards else None, float(discount),
                                 float(illegal_value), _ptr(q), n, _stream()))
    return q


def sample_children(ex: 'ExpandOut', k: int, *, seed: int = 0, env_offset: int = 0, draw_base: int = 0) -> torch.Tensor:
    """int64[N,4,k]: k chance children per legal afterstate DRAWN from the spawn distribution
    (g2048_sample_children; progressive widening over gameboard.py:174 / :205).  ``ex`` is any
    ``expand`` result (``children='none'`` is enough); illegal actions give zeros."""
    n = ex.after.shape[0]
    out = torch.empty((n, 4, int(k)), dtype=torch.int64, device=ex.after.device)
    with _on(ex.after.device):
        check(lib().g2048_sample_children(_ptr(ex.after), _ptr(ex.child_count), int(k), seed & (2**64 - 1),
                                          env_offset, draw_base & 0xFFFFFFFF, _ptr(out), n, _stream()))
    return out


def expect_sampled(child_values: torch.Tensor, ex: 'ExpandOut', k: int, discount: float = 1.0,
                   illegal_value: float = float('-inf'), use_rewards: bool = True) -> torch.Tensor:
    """float32[N,4]: reward + discount * mean of the k sampled child values (g2048_expect_sampled)."""
    n = ex.after.shape[0]
    _req(child_values, torch.float32, 'child_values', n * 4 * int(k))
    q = torch.empty((n, 4), dtype=torch.float32, device=ex.after.device)
    with _on(ex.after.device):
        check(lib().g2048_expect_sampled(_ptr(child_values), _ptr(ex.child_count),
                                         _ptr(ex.rewards) if use_rewards else None, int(k), float(discount),
                                         float(illegal_value), _ptr(q), n, _stream()))
    return q


def nstep_returns(rewards: torch.Tensor, values: torch.Tensor, discount: float, n_steps: int,
                  td_lambda: float = 0.5, lengths: Optional[torch.Tensor] = None,
                  out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """n-step TD(lambda) returns over time-major trajectory buffers, float32[T,N]
    (g2048_nstep_returns; self_play.py:524 compute_n_step_returns per env).  ``lengths`` int32[N]
    limits each env's trajectory; steps beyond it come back as 0."""
    if rewards.dim() != 2 or rewards.shape != values.shape:
        raise ValueError('rewards and values must both be [T, N]')
    for name, t in (('rewards', rewards), ('values', values)):
        if not t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
            raise ValueError(f'{name}: contiguous float32 CUDA tensor required')
    T, n = rewards.shape
    _opt(lengths, torch.int32, 'lengths', n)
    if out is None:
        out = torch.empty_like(rewards)
    _opt(out, torch.float32, 'out', T * n)
    with _on(rewards.device):
        check(lib().g2048_nstep_returns(_ptr(rewards), _ptr(values), _ptr(lengths), float(discount),
                                        int(n_steps), float(td_lambda), _ptr(out), T, n, _stream()))
    return out


class RolloutOut(NamedTuple):
    boards: torch.Tensor   # int64[N] final boards
    score: torch.Tensor    # int64[N]
    steps: torch.Tensor    # int32[N]
    maxexp: torch.Tensor   # uint8[N]
    stats: torch.Tensor    # int64[24] (see include/g2048.h)
    trace: Optional[torch.Tensor]


def rollout(n: int, *, seed: int = 0, env_offset: int = 0, max_steps: int = 1 << 20,
            boards: Optional[torch.Tensor] = None, trace_stride: int = 0,
            per_env: bool = True, stats: Optional[torch.Tensor] = None, device=None) -> RolloutOut:
    """Fused random-policy rollouts to termination (g2048_rollout).  ``boards=None`` starts from
    reset boards; ``per_env=False`` keeps only the statistics vector."""
    dev = torch.device('cuda' if device is None else device) if boards is None else boards.device
    from_reset = boards is None
    if boards is not None:
        boards = _req(boards, torch.int64, 'boards', n).clone()
    elif per_env:
        boards = torch.empty(n, dtype=torch.int64, device=dev)
    score = torch.empty(n, dtype=torch.int64, device=dev) if per_env else None
    steps = torch.empty(n, dtype=torch.int32, device=dev) if per_env else None
    mx = torch.empty(n, dtype=torch.uint8, device=dev) if per_env else None
    if stats is None:
        stats = torch.zeros(_cabi.STATS_WORDS, dtype=torch.int64, device=dev)
    else:
        _req(stats, torch.int64, 'stats', _cabi.STATS_WORDS)
    trace = torch.zeros((n, trace_stride), dtype=torch.uint8, device=dev) if trace_stride else None
    with _on(dev):
        check(lib().g2048_rollout(seed & (2**64 - 1), env_offset, n, max_steps & 0xFFFFFFFF,
                                  1 if from_reset else 0, _ptr(boards), _ptr(score), _ptr(steps),
                                  _ptr(mx), _ptr(stats), _ptr(trace), trace_stride, _stream()))
    return RolloutOut(boards, score, steps, mx, stats, trace)


def pack(tiles: torch.Tensor) -> torch.Tensor:
    """int32[N,4,4] (or [N,16]) tile values -> packed int64[N]; raises on unrepresentable tiles."""
    _req(tiles, torch.int32, 'tiles')
    if tiles.numel() % 16:
        raise ValueError('tiles: need a multiple of 16 elements')
    n = tiles.numel() // 16
    boards = torch.empty(n, dtype=torch.int64, device=tiles.device)
    status = torch.zeros(1, dtype=torch.int32, device=tiles.device)
    with _on(tiles.device):
        check(lib().g2048_pack(_ptr(tiles), _ptr(boards), _ptr(status), n, _stream()))
    code = int(status.item())
    if code:
        check(code)
    return boards


def unpack(boards: torch.Tensor) -> torch.Tensor:
    """packed int64[N] -> int32[N,4,4] tile values."""
    _req(boards, torch.int64, 'boards')
    n = boards.numel()
    tiles = torch.empty((n, 4, 4), dtype=torch.int32, device=boards.device)
    with _on(boards.device):
        check(lib().g2048_unpack(_ptr(boards), _ptr(tiles), n, _stream()))
    return tiles


_NP_DTYPES = {torch.int64: np.int64, torch.int32: np.int32, torch.uint8: np.uint8, torch.float32: np.float32}


def unpack_into(boards: torch.Tensor, tiles: torch.Tensor) -> None:
    """g2048_unpack into a preallocated int32[N,4,4] tensor."""
    _req(boards, torch.int64, 'boards')
    _req(tiles, torch.int32, 'tiles', boards.numel() * 16)
    with _on(boards.device):
        check(lib().g2048_unpack(_ptr(boards), _ptr(tiles), boards.numel(), _stream()))


def pinned_empty(n: int, dtype, device: Optional[int] = None) -> np.ndarray:
    """NumPy array backed by pinned host memory from ``g2048_host_alloc``: full PCIe bandwidth for
    ``step_host``, placed on the NUMA node of ``device`` (default: the current CUDA device) when the
    process is allowed to.  The block is returned to the library when the array and all its views
    are gone."""
    import ctypes
    import weakref
    np_dtype = np.dtype(_NP_DTYPES.get(dtype, dtype))
    nbytes = max(int(n) * np_dtype.itemsize, 1)
    p = ctypes.c_void_p()
    with torch.cuda.device(torch.cuda.current_device() if device is None else device):
        check(lib().g2048_host_alloc(ctypes.byref(p), nbytes))
    buf = (ctypes.c_uint8 * nbytes).from_address(p.value)
    weakref.finalize(buf, lib().g2048_host_free, ctypes.c_void_p(p.value))
    return np.frombuffer(buf, dtype=np_dtype, count=int(n))


def host_numa_info(arr: Optional[np.ndarray] = None):
    """(NUMA node of the current GPU, NUMA node holding ``arr``'s first page); -1 = unknown."""
    import ctypes
    g, m = ctypes.c_int(-1), ctypes.c_int(-1)
    check(lib().g2048_host_numa_info(None if arr is None else arr.ctypes.data, ctypes.byref(g), ctypes.byref(m)))
    return g.value, m.value


def shutdown() -> None:
    """Release the library's device resources (g2048_shutdown); it re-initialises lazily."""
    check(lib().g2048_shutdown())


def step_host(boards: np.ndarray, actions: np.ndarray, spawns: Optional[np.ndarray] = None, *,
              seed: int = 0, env_offset: int = 0, step_base: int = 0, out=None):
    """HOST ndarray in, HOST ndarray out (g2048_step_host): the reference-facing call.  The H2D
    copy, kernel and D2H copy of successive chunks overlap inside the library."""
    b = np.ascontiguousarray(boards).view(np.uint64).reshape(-1)
    n = b.size
    a = np.ascontiguousarray(actions, dtype=np.uint8).reshape(-1)
    s = None if spawns is None else np.ascontiguousarray(spawns, dtype=np.uint8).reshape(-1)
    if a.size != n or (s is not None and s.size != n):
        raise ValueError('actions / spawns must have one entry per board')
    if out is None:
        out = (np.empty(n, dtype=np.uint64), np.empty(n, dtype=np.int32), np.empty(n, dtype=np.uint8))
    ob, orw, ofl = out
    check(lib().g2048_step_host(b.ctypes.data, a.ctypes.data, None if s is None else s.ctypes.data,
                                seed & (2**64 - 1), env_offset, step_base & 0xFFFFFFFF,
                                ob.ctypes.data, orw.ctypes.data, ofl.ctypes.data, n))
    return ob, orw, ofl


def step_host_copies(boards: np.ndarray, actions: np.ndarray, spawns: Optional[np.ndarray], out) -> None:
    """The host<->device copies of ``step_host`` without the kernel (g2048_step_host_copies): the
    PCIe ceiling of that call, measured with the same buffers."""
    b = np.ascontiguousarray(boards).view(np.uint64).reshape(-1)
    ob, orw, ofl = out
    check(lib().g2048_step_host_copies(b.ctypes.data, actions.ctypes.data,
                                       None if spawns is None else spawns.ctypes.data,
                                       ob.ctypes.data, orw.ctypes.data, ofl.ctypes.data, b.size))


def row_table():
    """(result u16[65536], reward u32[65536], moved u8[65536]) read back from the device table."""
    res = np.empty(65536, dtype=np.uint16)
    rew = np.empty(65536, dtype=np.uint32)
    mv = np.empty(65536, dtype=np.uint8)
    check(lib().g2048_row_table_host(res.ctypes.data, rew.ctypes.data, mv.ctypes.data))
    return res, rew, mv
