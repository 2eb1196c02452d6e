"""GPU parity of the transition kernels (g2048_step / g2048_env_step / g2048_mask / g2048_spawn /
g2048_reset / g2048_observe / pack / unpack) against the pinned CPU oracle and the golden
fixtures of the real reference.  Everything goes through the C ABI.  Bar: bit-exact."""

import hashlib

import numpy as np
import pytest

from tests.helpers import pack_tiles, synthetic_packed, valid_spawn_bytes

pytestmark = pytest.mark.gpu

FLAG_DONE, FLAG_MOVED, FLAG_OVERFLOW, FLAG_BADSPAWN = 0x10, 0x20, 0x40, 0x80


@pytest.fixture(scope='module')
def T():
    import torch
    return torch


@pytest.fixture(scope='module')
def ops():
    from simulate_2048_b200 import ops
    ops.init()
    return ops


def dev(T, a, dtype=None):
    a = np.ascontiguousarray(a)
    if a.dtype == np.uint64:
        a = a.view(np.int64)
    t = T.from_numpy(a).cuda()
    return t if dtype is None else t.to(dtype)


def host(t):
    a = t.cpu().numpy()
    return a.view(np.uint64) if a.dtype == np.int64 else a


def run_step(T, ops, boards, actions, spawns=None, **kw):
    out = ops.step(dev(T, boards), dev(T, actions), None if spawns is None else dev(T, spawns), **kw)
    return host(out.boards), out.rewards.cpu().numpy(), out.flags.cpu().numpy()


def afterstates(oracle, boards, actions):
    ex = oracle.expand(boards)
    return ex['after'][np.arange(boards.size), actions.astype(np.int64) & 3]


# ------------------------------------------------------------------ row table

def test_device_row_table_hashes(ops, oracle):
    res, rew, mv = ops.row_table()
    o = oracle.row_table()
    np.testing.assert_array_equal(res, o[0])
    np.testing.assert_array_equal(rew, o[1])
    np.testing.assert_array_equal(mv, o[2])
    assert hashlib.sha256(res.tobytes()).hexdigest() == '2311124775fc20afc51de212e97fc3f04eef16691d5d270594080c77806f1cac'
    assert hashlib.sha256(rew.tobytes()).hexdigest() == '5998c5d66dcf2d4a7725a856533105ebffeff7fd3ef1825833cbbe9fc6bc9949'


# ------------------------------------------------------------------ golden fixtures of the reference

def test_step_reproduces_reference_next_state(T, ops, golden):
    g = golden('synthetic')
    legal = g['next_spawn'] != 0xFF
    spawns = np.where(legal, g['next_spawn'], 0).astype(np.uint8)
    out, rew, fl = run_step(T, ops, g['packed'], g['actions'], spawns)
    np.testing.assert_array_equal(out, g['next'])
    np.testing.assert_array_equal(rew, g['next_reward'].astype(np.int32))
    np.testing.assert_array_equal((fl & FLAG_MOVED) != 0, legal)
    assert not (fl & (FLAG_OVERFLOW | FLAG_BADSPAWN)).any()


def test_step_replays_reference_games(T, ops, golden):
    g = golden('games')
    legal = g['spawn'] != 0xFF
    spawns = np.where(legal, g['spawn'], 0).astype(np.uint8)
    out, rew, fl = run_step(T, ops, g['board'], g['action'], spawns)
    np.testing.assert_array_equal(out, g['next'])
    np.testing.assert_array_equal(rew, g['reward'].astype(np.int32))
    np.testing.assert_array_equal((fl & FLAG_DONE) != 0, g['done'] != 0)


def test_mask_matches_reference(T, ops, golden):
    g = golden('synthetic')
    fl = ops.legal_flags(dev(T, g['packed'])).cpu().numpy()
    np.testing.assert_array_equal(fl & 0xF, g['mask'])
    np.testing.assert_array_equal((fl & FLAG_DONE) != 0, g['done'] != 0)


# ------------------------------------------------------------------ oracle parity, every code path

@pytest.mark.parametrize('impl', [0, 1, 2])
@pytest.mark.parametrize('n', [1, 5, 4096, 4096 * 3 + 17, (1 << 20) + 37])
def test_step_injected_vs_oracle(T, ops, oracle, n, impl):
    """impl 0 = auto (general kernel below 2^18 envs, smem-table kernel above), 1 = smem table
    forced (covers its ragged tail at every size), 2 = pure register arithmetic (no table)."""
    boards = synthetic_packed(n, seed=100 + n % 7)
    rng = np.random.default_rng(n)
    actions = rng.integers(0, 4, n).astype(np.uint8)
    spawns = valid_spawn_bytes(afterstates(oracle, boards, actions), seed=n)
    ops.set_impl(impl)
    try:
        out, rew, fl = run_step(T, ops, boards, actions, spawns)
    finally:
        ops.set_impl(0)
    o_out, o_rew, o_fl = oracle.step(boards, actions, spawns)
    np.testing.assert_array_equal(out, o_out)
    np.testing.assert_array_equal(rew, o_rew)
    np.testing.assert_array_equal(fl, o_fl)


@pytest.mark.parametrize('impl', [0, 1, 2])
@pytest.mark.parametrize('n', [3, 4096 + 5, (1 << 19) + 3])
def test_step_philox_vs_oracle(T, ops, oracle, n, impl):
    boards = synthetic_packed(n, seed=55)
    actions = np.random.default_rng(n).integers(0, 4, n).astype(np.uint8)
    kw = dict(seed=0xC0FFEE1234567, env_offset=(1 << 33) + 5, step_base=77)
    ops.set_impl(impl)
    try:
        out, rew, fl = run_step(T, ops, boards, actions, None, **kw)
    finally:
        ops.set_impl(0)
    o_out, o_rew, o_fl = oracle.step(boards, actions, None, **kw)
    np.testing.assert_array_equal(out, o_out)
    np.testing.assert_array_equal(rew, o_rew)
    np.testing.assert_array_equal(fl, o_fl)


@pytest.mark.parametrize('threads,envs', [(512, 4), (640, 4), (768, 4), (1024, 4), (512, 8), (1024, 8)])
def test_step_launch_geometries_vs_oracle(T, ops, oracle, threads, envs):
    """Every instantiated CTA size / envs-per-thread of the smem-table kernel, both spawn modes,
    at a size that leaves a ragged tail for each tile size."""
    n = (1 << 19) + 4099
    boards = synthetic_packed(n, seed=threads + envs)
    actions = np.random.default_rng(threads).integers(0, 4, n).astype(np.uint8)
    spawns = valid_spawn_bytes(afterstates(oracle, boards, actions), seed=envs)
    ops.set_tuning(threads, envs)
    try:
        got_inj = run_step(T, ops, boards, actions, spawns)
        got_phx = run_step(T, ops, boards, actions, None, seed=99, step_base=3)
    finally:
        ops.set_tuning(0, 4)
    for got, want in ((got_inj, oracle.step(boards, actions, spawns)),
                      (got_phx, oracle.step(boards, actions, None, seed=99, step_base=3))):
        for g, w in zip(got, want):
            np.testing.assert_array_equal(g, w)
    with pytest.raises(Exception):
        ops.set_tuning(640, 8)


def test_step_optional_inputs_vs_oracle(T, ops, oracle):
    n = 10007
    boards = synthetic_packed(n, seed=8)
    rng = np.random.default_rng(8)
    actions = rng.integers(0, 4, n).astype(np.uint8)
    step_ctr = rng.integers(0, 1 << 31, n).astype(np.uint32)
    keys = rng.integers(0, 1 << 63, n).astype(np.uint64)
    done_in = (rng.random(n) < 0.2).astype(np.uint8)
    out = ops.step(dev(T, boards), dev(T, actions), None, step_ctr=dev(T, step_ctr.view(np.int32)),
                   env_keys=dev(T, keys), done_in=dev(T, done_in), env_offset=3)
    o_out, o_rew, o_fl = oracle.step(boards, actions, None, step_ctr=step_ctr, env_keys=keys,
                                     done_in=done_in, env_offset=3)
    np.testing.assert_array_equal(host(out.boards), o_out)
    np.testing.assert_array_equal(out.rewards.cpu().numpy(), o_rew)
    np.testing.assert_array_equal(out.flags.cpu().numpy(), o_fl)
    passed = done_in != 0
    np.testing.assert_array_equal(o_out[passed], boards[passed])
    assert (o_rew[passed] == 0).all() and ((o_fl[passed] & FLAG_DONE) != 0).all()


def test_step_big_tiles_overflow_and_slow_path(T, ops, oracle):
    """Exponents up to 15: rows outside the staged table window and 15+15 merges must be exact,
    in the general, the smem-table and the register-arithmetic kernel."""
    n = (1 << 18) + 11
    boards = synthetic_packed(n, seed=9, max_exp=15)
    # force plenty of 15+15 pairs
    boards[::97] = np.uint64(0x00FF00000FF0FF00)
    actions = np.random.default_rng(1).integers(0, 4, n).astype(np.uint8)
    spawns = valid_spawn_bytes(afterstates(oracle, boards, actions), seed=4)
    o = oracle.step(boards, actions, spawns)
    assert (o[2] & FLAG_OVERFLOW).any()
    for impl in (0, 1, 2):
        ops.set_impl(impl)
        try:
            got = run_step(T, ops, boards, actions, spawns)
        finally:
            ops.set_impl(0)
        for a, b in zip(got, o):
            np.testing.assert_array_equal(a, b)


def test_step_spawn_byte_forms(T, ops, oracle):
    n = 5000
    boards = synthetic_packed(n, seed=21)
    rng = np.random.default_rng(21)
    actions = rng.integers(0, 4, n).astype(np.uint8)
    # arbitrary bytes: absolute cells (bit 7), out-of-range ranks, occupied cells
    spawns = rng.integers(0, 256, n).astype(np.uint8)
    got = run_step(T, ops, boards, actions, spawns)
    want = oracle.step(boards, actions, spawns)
    for a, b in zip(got, want):
        np.testing.assert_array_equal(a, b)
    assert (want[2] & FLAG_BADSPAWN).any()


def test_step_edge_boards(T, ops, oracle):
    tiles = np.array([
        [[0] * 4] * 4,                                                        # empty: nothing legal
        [[2, 4, 2, 4], [4, 2, 4, 2], [2, 4, 2, 4], [4, 2, 4, 2]],             # checkerboard: done
        [[2, 0, 0, 0], [4, 0, 0, 0], [8, 0, 0, 0], [16, 0, 0, 0]],            # left illegal
        [[2, 2, 2, 2], [2, 2, 2, 2], [2, 2, 2, 2], [2, 2, 2, 2]],             # full, everything merges
        [[32768, 16384, 8192, 4096], [2048, 1024, 512, 256], [128, 64, 32, 16], [8, 4, 2, 0]],
    ])
    boards = np.repeat(pack_tiles(tiles), 4)
    actions = np.tile(np.arange(4, dtype=np.uint8), len(tiles))
    spawns = np.zeros(boards.size, np.uint8)
    got = run_step(T, ops, boards, actions, spawns)
    want = oracle.step(boards, actions, spawns)
    for a, b in zip(got, want):
        np.testing.assert_array_equal(a, b)
    assert (got[2][:4] == FLAG_DONE).all() and (got[1][:8] == 0).all()
    # n = 0 is a no-op
    empty = ops.step(dev(T, boards[:0]), dev(T, actions[:0]), dev(T, spawns[:0]))
    assert empty.boards.numel() == 0


def test_step_in_place_aliasing(T, ops, oracle):
    from simulate_2048_b200.ops import StepOut
    n = (1 << 18) + 9
    boards = synthetic_packed(n, seed=2)
    actions = np.random.default_rng(2).integers(0, 4, n).astype(np.uint8)
    d = dev(T, boards)
    out = StepOut(d, T.empty(n, dtype=T.int32, device='cuda'), T.empty(n, dtype=T.uint8, device='cuda'))
    ops.step(d, dev(T, actions), None, seed=4, out=out)
    np.testing.assert_array_equal(host(d), oracle.step(boards, actions, None, seed=4)[0])


# ------------------------------------------------------------------ full-size properties (config 2)

def test_step_full_size_properties(T, ops, oracle):
    """2^24 boards: bit-exact against the C oracle (it finishes in seconds) and size-independent
    invariants: tile sum conserved up to the spawned tile, illegal => untouched, idempotence of
    the flags (mask kernel on the output equals the step's flags)."""
    n = 1 << 24
    boards = np.tile(synthetic_packed(1 << 20, seed=2048), 16)  # the 1M config-2 set, 16 times
    actions = np.random.default_rng(0).integers(0, 4, n).astype(np.uint8)
    d_b, d_a = dev(T, boards), dev(T, actions)
    out = ops.step(d_b, d_a, None, seed=2048)
    got_b, got_r, got_f = host(out.boards), out.rewards.cpu().numpy(), out.flags.cpu().numpy()
    o_b, o_r, o_f = oracle.step(boards, actions, None, seed=2048, threads=8)
    assert (got_b == o_b).all() and (got_r == o_r).all() and (got_f == o_f).all()
    np.testing.assert_array_equal(ops.legal_flags(out.boards).cpu().numpy(), got_f & 0x1F)
    moved = (got_f & FLAG_MOVED) != 0
    assert (got_b[~moved] == boards[~moved]).all() and (got_r[~moved] == 0).all()

    def tile_sum(b):
        s = np.zeros(b.size, np.int64)
        for i in range(16):
            e = ((b >> np.uint64(4 * i)) & np.uint64(0xF)).astype(np.int64)
            s += np.where(e > 0, np.int64(1) << e, 0)
        return s
    delta = tile_sum(got_b) - tile_sum(boards)
    assert set(np.unique(delta[moved])) <= {2, 4} and (delta[~moved] == 0).all()
    assert abs((delta[moved] == 4).mean() - 0.1) < 0.002


# ------------------------------------------------------------------ env_step (config 5 semantics)

def test_env_step_trajectories_vs_oracle(T, ops, oracle):
    """4096 envs x 200 steps (config.py:51,79 scaled to config 5), injected spawns, including
    illegal actions and envs that finish early and must pass through untouched."""
    n, steps = 4096, 200
    rng = np.random.default_rng(5)
    b0 = oracle.reset(n, seed=1)
    st = (b0, np.zeros(n, np.int32), np.zeros(n, np.uint8), np.zeros(n, np.float32))
    d_b = dev(T, b0)
    d_sc = T.zeros(n, dtype=T.int32, device='cuda')
    d_done = T.zeros(n, dtype=T.uint8, device='cuda')
    d_tr = T.zeros(n, dtype=T.float32, device='cuda')
    d_rew = T.empty(n, dtype=T.float32, device='cuda')
    d_fl = T.empty(n, dtype=T.uint8, device='cuda')
    d_obs = T.empty((n, 16), dtype=T.float32, device='cuda')
    d_mt = T.empty(n, dtype=T.int32, device='cuda')
    for t in range(steps):
        actions = rng.integers(0, 4, n).astype(np.uint8)
        spawns = valid_spawn_bytes(afterstates(oracle, st[0], actions), seed=t)
        st, rew, fl, obs, mt = oracle.env_step(*st, actions, spawns)
        ops.env_step(d_b, d_sc, d_done, d_tr, dev(T, actions), dev(T, spawns), rewards=d_rew,
                     flags=d_fl, obs=d_obs, max_tile=d_mt)
        if t % 20 == 19 or t == steps - 1:
            np.testing.assert_array_equal(host(d_b), st[0])
            np.testing.assert_array_equal(d_sc.cpu().numpy(), st[1])
            np.testing.assert_array_equal(d_done.cpu().numpy(), st[2])
            np.testing.assert_array_equal(d_tr.cpu().numpy(), st[3])
            np.testing.assert_array_equal(d_rew.cpu().numpy(), rew)
            np.testing.assert_array_equal(d_fl.cpu().numpy(), fl)
            np.testing.assert_array_equal(d_obs.cpu().numpy(), obs)
            np.testing.assert_array_equal(d_mt.cpu().numpy(), mt)
    assert st[2].sum() > 0  # some games did finish inside the horizon


def test_env_step_philox_keys(T, ops, oracle):
    n = 3000
    b0 = oracle.reset(n, seed=4)
    keys = np.random.default_rng(4).integers(0, 1 << 63, n).astype(np.uint64)
    st = (b0, np.zeros(n, np.int32), np.zeros(n, np.uint8), np.zeros(n, np.float32))
    d = [dev(T, b0), T.zeros(n, dtype=T.int32, device='cuda'), T.zeros(n, dtype=T.uint8, device='cuda'),
         T.zeros(n, dtype=T.float32, device='cuda')]
    rng = np.random.default_rng(6)
    for _ in range(50):
        actions = rng.integers(0, 4, n).astype(np.uint8)
        st, *_ = oracle.env_step(*st, actions, None, env_keys=keys)
        ops.env_step(*d, dev(T, actions), None, env_keys=dev(T, keys))
    np.testing.assert_array_equal(host(d[0]), st[0])
    np.testing.assert_array_equal(d[3].cpu().numpy(), st[3])


# ------------------------------------------------------------------ reset / spawn / observe / pack

def test_reset_spawn_observe_pack_vs_oracle(T, ops, oracle, golden):
    n = 50000
    b = host(ops.reset(n, seed=77, env_offset=123456789012))
    np.testing.assert_array_equal(b, oracle.reset(n, seed=77, env_offset=123456789012))
    sp = np.random.default_rng(0).integers(0, 16, 2 * n).astype(np.uint8)
    np.testing.assert_array_equal(host(ops.reset(n, spawns=dev(T, sp))), oracle.reset(n, spawns=sp))
    keys = np.random.default_rng(1).integers(0, 1 << 63, n).astype(np.uint64)
    np.testing.assert_array_equal(host(ops.reset(n, env_keys=dev(T, keys))), oracle.reset(n, env_keys=keys))
    # reference properties (tests/test_environment.py:31-46, :249-279)
    tiles = oracle.unpack(b)
    assert ((tiles != 0).sum(axis=(1, 2)) == 2).all() and set(np.unique(tiles)) == {0, 2, 4}
    assert abs((tiles[tiles != 0] == 4).mean() - 0.1) < 0.01

    boards = synthetic_packed(n, seed=31)
    got = ops.spawn(dev(T, boards), None, seed=5, step=9, tag=3)
    np.testing.assert_array_equal(host(got), oracle.spawn(boards, None, seed=5, step=9, tag=3)[0])
    bytes_ = np.random.default_rng(2).integers(0, 256, n).astype(np.uint8)
    flags = T.empty(n, dtype=T.uint8, device='cuda')
    got = ops.spawn(dev(T, boards), dev(T, bytes_), flags=flags)
    want_b, want_f = oracle.spawn(boards, bytes_)
    np.testing.assert_array_equal(host(got), want_b)
    np.testing.assert_array_equal(flags.cpu().numpy(), want_f)
    # fill_cells draws of the reference (fixture) through absolute-cell bytes
    g = golden('fill_cells')
    base = np.repeat(np.uint64(g['base']), len(g['draws']))
    by = np.array([r | (0x10 if v == 4 else 0) for _, v, r in g['draws'].tolist()], np.uint8)
    got = host(ops.spawn(dev(T, base), dev(T, by)))
    want = np.array([int(g['base']) | ((2 if v == 4 else 1) << (4 * c)) for c, v, _ in g['draws'].tolist()], np.uint64)
    np.testing.assert_array_equal(got, want)

    obs, mx = ops.observe(dev(T, boards))
    o_obs, o_mx = oracle.observe(boards)
    np.testing.assert_array_equal(obs.cpu().numpy(), o_obs)
    np.testing.assert_array_equal(mx.cpu().numpy(), o_mx)

    tiles = oracle.unpack(boards)
    np.testing.assert_array_equal(ops.unpack(dev(T, boards)).cpu().numpy(), tiles)
    np.testing.assert_array_equal(host(ops.pack(dev(T, tiles))), boards)
    bad = tiles.copy()
    bad[7, 1, 1] = 3
    from simulate_2048_b200._cabi import G2048Error
    with pytest.raises(G2048Error):
        ops.pack(dev(T, bad))
    bad[7, 1, 1] = 65536
    with pytest.raises(G2048Error):
        ops.pack(dev(T, bad))


def test_supporting_kernels_ragged_and_unaligned(T, ops, oracle):
    """The wide-access kernels (mask, observe, one-hot, pack, unpack) on batch sizes that are not
    multiples of their 4-board / 32-board tiles, and through pointers that only have the natural
    alignment of their element type (scalar fall-backs of g2048_mask / pack / unpack / one-hot)."""
    from simulate_2048_b200 import _cabi
    lib = ops.lib()
    for n in (1, 3, 31, 33, 127, 4099):
        boards = synthetic_packed(n + 1, seed=100 + n)
        boards[0] = 0xFEDCBA9876543210
        d = dev(T, boards)
        want_t = oracle.unpack(boards)
        want_f = oracle.mask(boards) if hasattr(oracle, 'mask') else None
        o_obs, o_mx = oracle.observe(boards)
        for off in (0, 1):      # off = 1: boards pointer 8-byte aligned only
            v = d[off:]
            m = v.numel()
            obs, mx = ops.observe(v.contiguous())
            np.testing.assert_array_equal(obs.cpu().numpy(), o_obs[off:])
            np.testing.assert_array_equal(mx.cpu().numpy(), o_mx[off:])
            flags = T.full((m + 8,), 0xEE, dtype=T.uint8, device='cuda')
            fv = flags[1 + off:1 + off + m]          # flags pointer 1-byte aligned
            assert lib.g2048_mask(v.data_ptr(), fv.data_ptr(), m, None) == 0
            ref = ops.legal_flags(v.contiguous())
            T.cuda.synchronize()
            np.testing.assert_array_equal(fv.cpu().numpy(), ref.cpu().numpy())
            assert int(flags[off].item()) == 0xEE and int(flags[1 + off + m].item()) == 0xEE
            if want_f is not None:
                np.testing.assert_array_equal(ref.cpu().numpy(), want_f[off:])
            tiles = T.full((m * 16 + 8,), -7, dtype=T.int32, device='cuda')
            tv = tiles[1:1 + m * 16]                 # tiles pointer 4-byte aligned only
            assert lib.g2048_unpack(v.data_ptr(), tv.data_ptr(), m, None) == 0
            T.cuda.synchronize()
            np.testing.assert_array_equal(tv.cpu().numpy().reshape(-1, 4, 4), want_t[off:])
            assert int(tiles[0].item()) == -7 and int(tiles[1 + m * 16].item()) == -7
            packed = T.zeros(m, dtype=T.int64, device='cuda')
            assert lib.g2048_pack(tv.data_ptr(), packed.data_ptr(), None, m, None) == 0
            T.cuda.synchronize()
            np.testing.assert_array_equal(host(packed), boards[off:])
            np.testing.assert_array_equal(host(ops.pack(tv.clone().reshape(m, 4, 4))), boards[off:])
            for width in (16, 31, 64):
                oh = ops.observe_onehot(v.contiguous(), width).cpu().numpy().reshape(m, 16, width)
                exps = np.log2(np.maximum(want_t[off:].reshape(m, 16), 1)).astype(np.int64)
                np.testing.assert_array_equal(oh, np.eye(width, dtype=np.uint8)[exps])
                raw = T.full((m * 16 * width + 32,), 9, dtype=T.uint8, device='cuda')
                rv = raw[3:3 + m * 16 * width]       # 1-byte aligned output: byte kernel
                assert lib.g2048_observe_onehot(v.data_ptr(), rv.data_ptr(), width, m, None) == 0
                T.cuda.synchronize()
                np.testing.assert_array_equal(rv.cpu().numpy().reshape(m, 16, width), oh)
                assert int(raw[2].item()) == 9 and int(raw[3 + m * 16 * width].item()) == 9


def test_host_buffer_step_matches_device_path(T, ops, oracle):
    """g2048_step_host (host ndarrays in and out, chunked copy/compute pipeline) == oracle."""
    n = (1 << 23) + 12345  # three chunks, the last one ragged
    boards = synthetic_packed(n, seed=13)
    actions = np.random.default_rng(13).integers(0, 4, n).astype(np.uint8)
    ob, orw, ofl = ops.step_host(boards, actions, None, seed=99, env_offset=7)
    want = oracle.step(boards, actions, None, seed=99, env_offset=7, threads=8)
    np.testing.assert_array_equal(ob, want[0])
    np.testing.assert_array_equal(orw, want[1])
    np.testing.assert_array_equal(ofl, want[2])
    spawns = valid_spawn_bytes(afterstates(oracle, boards[:100000], actions[:100000]))
    got = ops.step_host(boards[:100000], actions[:100000], spawns)
    want = oracle.step(boards[:100000], actions[:100000], spawns)
    for a, b in zip(got, want):
        np.testing.assert_array_equal(a, b)


def test_no_cpu_fallback(T, ops):
    with pytest.raises(ValueError):
        ops.step(T.zeros(4, dtype=T.int64), T.zeros(4, dtype=T.uint8))
