"""ctypes binding of oracle/g2048_oracle.c (test infrastructure; see that file's header).

Every function takes/returns NumPy arrays on the host and mirrors the signature of the
matching ``g2048_*`` entry point in include/g2048.h, so parity tests read
``assert_array_equal(cuda(...), oracle(...))``.
"""

from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, 'libg2048_oracle.so')
_lib = None

STATS_WORDS = 24


def build(force: bool = False) -> str:
    """Compile the C oracle with gcc (a few hundred ms)."""
    src = os.path.join(_HERE, 'g2048_oracle.c')
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(
            ['gcc', '-O2', '-fPIC', '-pthread', '-std=c11', '-shared', '-o', _SO, src]
        )
    return _SO


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.orc_latent_state.restype = C.c_int64
        _lib.orc_latent_state_values.restype = C.c_int64
        _lib.orc_merge_line_values.restype = C.c_int64
        _lib.orc_legal_mask.restype = C.c_uint
        _lib.orc_legal_mask_values.restype = C.c_uint
        _lib.orc_rollout.restype = C.c_uint64
        _lib.orc_philox_word.restype = C.c_uint32
        _lib.orc_game_word.restype = C.c_uint32
        _lib.orc_draw_spawn.restype = C.c_uint32
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _u64(x):
    return np.ascontiguousarray(x, dtype=np.uint64)


def _u8(x):
    return None if x is None else np.ascontiguousarray(x, dtype=np.uint8)


def philox4x32_10(ctr, key):
    c = np.ascontiguousarray(ctr, dtype=np.uint32)
    k = np.ascontiguousarray(key, dtype=np.uint32)
    out = np.zeros(4, dtype=np.uint32)
    lib().orc_philox4x32_10(_p(c), _p(k), _p(out))
    return out


def philox_word(key: int, env: int, step: int, tag: int) -> int:
    """The 32-bit word env reads at (step, tag): word env & 3 of the block of quad env >> 2."""
    return int(lib().orc_philox_word(C.c_uint64(key), C.c_uint64(env), C.c_uint32(step), C.c_uint32(tag)))


def game_word(key: int, env: int, p: int) -> int:
    """Word p of env's game stream (g2048_reset, g2048_rollout): word p & 3 of the env's block p >> 2."""
    return int(lib().orc_game_word(C.c_uint64(key), C.c_uint64(env), C.c_uint32(p)))


def draw_spawn(w: int, n_empty: int):
    """(rank, four, remainder): the first two mixed-radix digits of a word."""
    rank, four = C.c_uint32(0), C.c_uint32(0)
    rem = lib().orc_draw_spawn(C.c_uint32(w), C.c_uint32(n_empty), C.byref(rank), C.byref(four))
    return int(rank.value), bool(four.value), int(rem)


def sample_children(after, child_count, k, seed=0, env_offset=0, draw_base=0):
    a = _u64(after).reshape(-1)
    cc = _u8(child_count).reshape(-1)
    out = np.zeros(a.size * k, dtype=np.uint64)
    lib().orc_sample_children(_p(a), _p(cc), C.c_uint32(k), C.c_uint64(seed), C.c_uint64(env_offset),
                              C.c_uint32(draw_base), _p(out), C.c_size_t(a.size // 4))
    return out.reshape(-1, 4, k)


def latent_state(board: int, action: int):
    after = C.c_uint64(0)
    ovf = C.c_int(0)
    r = lib().orc_latent_state(C.c_uint64(int(board)), C.c_int(action), C.byref(after), C.byref(ovf))
    return int(after.value), int(r), bool(ovf.value)


def latent_state_values(tiles, action: int):
    t = np.ascontiguousarray(tiles, dtype=np.int64).reshape(16)
    out = np.zeros(16, dtype=np.int64)
    r = lib().orc_latent_state_values(_p(t), C.c_int(action), _p(out))
    return out.reshape(4, 4), int(r)


def merge_line_values(line):
    t = np.ascontiguousarray(line, dtype=np.int64).reshape(4)
    out = np.zeros(4, dtype=np.int64)
    r = lib().orc_merge_line_values(_p(t), _p(out))
    return out, int(r)


def legal_mask_values(tiles) -> int:
    t = np.ascontiguousarray(tiles, dtype=np.int64).reshape(16)
    return int(lib().orc_legal_mask_values(_p(t)))


def is_done_numpy_values(tiles) -> bool:
    t = np.ascontiguousarray(tiles, dtype=np.int64).reshape(16)
    return bool(lib().orc_is_done_numpy_values(_p(t)))


def step(boards, actions, spawns=None, seed=0, env_offset=0, step_ctr=None, step_base=0,
         env_keys=None, done_in=None, threads=None):
    b = _u64(boards)
    n = b.size
    a = _u8(actions)
    s = _u8(spawns)
    sc = None if step_ctr is None else np.ascontiguousarray(step_ctr, dtype=np.uint32)
    ek = None if env_keys is None else _u64(env_keys)
    di = _u8(done_in)
    out = np.empty(n, dtype=np.uint64)
    rew = np.empty(n, dtype=np.int32)
    fl = np.empty(n, dtype=np.uint8)
    if threads and threads > 1 and sc is None and ek is None and di is None:
        lib().orc_set_num_threads(int(threads))
        lib().orc_step_mt(_p(b), _p(a), _p(s), C.c_uint64(seed), C.c_uint64(env_offset),
                          C.c_uint32(step_base), _p(out), _p(rew), _p(fl), C.c_size_t(n))
    else:
        lib().orc_step(_p(b), _p(a), _p(s), C.c_uint64(seed), C.c_uint64(env_offset), _p(sc),
                       C.c_uint32(step_base), _p(ek), _p(di), _p(out), _p(rew), _p(fl),
                       C.c_size_t(n))
    return out, rew, fl


def env_step(boards, step_count, done, total_reward, actions, spawns=None, seed=0, env_offset=0,
             env_keys=None, want_obs=True):
    """Returns new (boards, step_count, done, total_reward), rewards, flags, obs, max_tile."""
    b = _u64(boards).copy()
    n = b.size
    sc = np.ascontiguousarray(step_count, dtype=np.int32).copy()
    d = _u8(done).copy()
    tr = np.ascontiguousarray(total_reward, dtype=np.float32).copy()
    a = _u8(actions)
    s = _u8(spawns)
    ek = None if env_keys is None else _u64(env_keys)
    rew = np.empty(n, dtype=np.float32)
    fl = np.empty(n, dtype=np.uint8)
    obs = np.empty((n, 16), dtype=np.float32) if want_obs else None
    mt = np.empty(n, dtype=np.int32)
    lib().orc_env_step(_p(b), _p(sc), _p(d), _p(tr), _p(a), _p(s), C.c_uint64(seed),
                       C.c_uint64(env_offset), _p(ek), _p(rew), _p(fl), _p(obs), _p(mt),
                       C.c_size_t(n))
    return (b, sc, d, tr), rew, fl, obs, mt


def mask(boards):
    b = _u64(boards)
    fl = np.empty(b.size, dtype=np.uint8)
    lib().orc_mask(_p(b), _p(fl), C.c_size_t(b.size))
    return fl


def expand(boards, child_layout=0):
    """Returns dict(after, rewards, mask, num_empty, child_count, children, child_offsets)."""
    b = _u64(boards)
    n = b.size
    after = np.empty((n, 4), dtype=np.uint64)
    rew = np.empty((n, 4), dtype=np.int32)
    m = np.empty(n, dtype=np.uint8)
    ne = np.empty((n, 4), dtype=np.uint8)
    cc = np.empty((n, 4), dtype=np.uint8)
    lib().orc_expand(_p(b), _p(after), _p(rew), _p(m), _p(ne), _p(cc), None, None, C.c_int(0),
                     C.c_size_t(n))
    out = dict(after=after, rewards=rew, mask=m, num_empty=ne, child_count=cc, children=None,
               child_offsets=None)
    if child_layout == 1:
        ch = np.zeros((n, 4, 16, 2), dtype=np.uint64)
        lib().orc_expand(_p(b), None, None, None, None, None, _p(ch), None, C.c_int(1), C.c_size_t(n))
        out['children'] = ch
    elif child_layout == 2:
        per_env = cc.astype(np.int64).sum(axis=1)
        off = np.zeros(n, dtype=np.int64)
        np.cumsum(per_env[:-1], out=off[1:])
        ch = np.zeros(int(per_env.sum()), dtype=np.uint64)
        lib().orc_expand(_p(b), None, None, None, None, None, _p(ch), _p(off), C.c_int(2),
                         C.c_size_t(n))
        out['children'] = ch
        out['child_offsets'] = off
    return out


def after_state(boards):
    b = _u64(boards)
    ch = np.zeros((b.size, 32), dtype=np.uint64)
    ne = np.empty(b.size, dtype=np.uint8)
    ec = np.empty(b.size, dtype=np.uint16)
    lib().orc_after_state(_p(b), _p(ch), _p(ne), _p(ec), C.c_size_t(b.size))
    return ch, ne, ec


def reset(n, seed=0, env_offset=0, env_keys=None, spawns=None):
    ek = None if env_keys is None else _u64(env_keys)
    s = _u8(spawns)
    out = np.empty(n, dtype=np.uint64)
    lib().orc_reset(C.c_uint64(seed), C.c_uint64(env_offset), _p(ek), _p(s), _p(out), C.c_size_t(n))
    return out


def spawn(boards, spawns=None, seed=0, env_offset=0, env_keys=None, step=0, tag=0):
    b = _u64(boards).copy()
    s = _u8(spawns)
    ek = None if env_keys is None else _u64(env_keys)
    fl = np.empty(b.size, dtype=np.uint8)
    lib().orc_spawn(_p(b), _p(s), C.c_uint64(seed), C.c_uint64(env_offset), _p(ek), C.c_uint32(step),
                    C.c_uint32(tag), _p(fl), C.c_size_t(b.size))
    return b, fl


def observe(boards):
    b = _u64(boards)
    obs = np.empty((b.size, 16), dtype=np.float32)
    mx = np.empty(b.size, dtype=np.uint8)
    lib().orc_observe(_p(b), _p(obs), _p(mx), C.c_size_t(b.size))
    return obs, mx


def pack(tiles):
    t = np.ascontiguousarray(tiles, dtype=np.int32).reshape(-1, 16)
    out = np.empty(t.shape[0], dtype=np.uint64)
    rc = lib().orc_pack(_p(t), _p(out), C.c_size_t(t.shape[0]))
    return out, int(rc)


def unpack(boards):
    b = _u64(boards)
    t = np.empty((b.size, 16), dtype=np.int32)
    lib().orc_unpack(_p(b), _p(t), C.c_size_t(b.size))
    return t.reshape(-1, 4, 4)


def rollout(n, seed=0, env_offset=0, max_steps=1 << 20, from_reset=True, boards=None,
            trace_stride=0):
    b = np.zeros(n, dtype=np.uint64) if boards is None else _u64(boards).copy()
    score = np.empty(n, dtype=np.int64)
    steps = np.empty(n, dtype=np.int32)
    mx = np.empty(n, dtype=np.uint8)
    stats = np.zeros(STATS_WORDS, dtype=np.uint64)
    trace = np.zeros((n, trace_stride), dtype=np.uint8) if trace_stride else None
    lib().orc_rollout(C.c_uint64(seed), C.c_uint64(env_offset), C.c_size_t(n), C.c_uint32(max_steps),
                      C.c_int(1 if from_reset else 0), _p(b), _p(score), _p(steps), _p(mx),
                      _p(stats), _p(trace), C.c_uint32(trace_stride))
    return dict(boards=b, score=score, steps=steps, maxexp=mx, stats=stats, trace=trace)


def rollout_stats_mt(n, seed=0, env_offset=0, max_steps=1 << 20, threads=1):
    stats = np.zeros(STATS_WORDS, dtype=np.uint64)
    lib().orc_set_num_threads(int(threads))
    lib().orc_rollout_mt(C.c_uint64(seed), C.c_uint64(env_offset), C.c_size_t(n),
                         C.c_uint32(max_steps), _p(stats))
    return stats


def row_table():
    res = np.empty(65536, dtype=np.uint16)
    rew = np.empty(65536, dtype=np.uint32)
    mv = np.empty(65536, dtype=np.uint8)
    lib().orc_row_table(_p(res), _p(rew), _p(mv))
    return res, rew, mv
