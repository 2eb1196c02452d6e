"""Library state and re-entrancy (include/g2048.h "State"): concurrent callers on several host
threads and streams, in-place stepping, shutdown and lazy re-initialisation, pinned host blocks;
plus the two mirrors the round-1 review found untested (game.Environment, create_initial_board)."""

import threading

import numpy as np
import pytest

from tests.helpers import synthetic_packed, valid_spawn_bytes

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def T():
    import torch
    return torch


@pytest.fixture(scope='module')
def ops():
    from simulate_2048_b200 import ops
    ops.init()
    return ops


def dev(T, a):
    return T.from_numpy(np.ascontiguousarray(a).view(np.int64)).cuda()


def host(t):
    return t.cpu().numpy().view(np.uint64)


def test_concurrent_threads_and_streams(T, ops, oracle):
    """Four host threads, each on its own stream, call g2048_step (both kernels), g2048_expand_compact
    and g2048_step_host at once; every result equals the oracle's.  Exercises the locked first-use
    initialisation (after a shutdown), the pipe pool and the per-call scratch allocations."""
    ops.shutdown()     # the first calls below race into current_state()
    n_big, n_small = (1 << 18) + 4 * 37, 5003
    jobs = []
    for k in range(4):
        b = synthetic_packed(n_big if k % 2 == 0 else n_small, seed=50 + k)
        a = np.random.default_rng(k).integers(0, 4, b.size).astype(np.uint8)
        jobs.append((b, a))
    want = [oracle.step(b, a, None, seed=11 + k, env_offset=4 * k) for k, (b, a) in enumerate(jobs)]
    want_x = [oracle.expand(b[:3000], child_layout=2) for b, _ in jobs]
    errors, results = [], [None] * 4

    def worker(k):
        try:
            b, a = jobs[k]
            stream = T.cuda.Stream()
            with T.cuda.stream(stream):
                d_b, d_a = dev(T, b), T.from_numpy(a).cuda()
                for _ in range(5):
                    out = ops.step(d_b, d_a, None, seed=11 + k, env_offset=4 * k)
                    ex = ops.expand(d_b[:3000].contiguous(), children='compact')
                hb, hr, hf = ops.step_host(b, a, None, seed=11 + k, env_offset=4 * k)
                stream.synchronize()
                results[k] = (host(out.boards), out.rewards.cpu().numpy(), out.flags.cpu().numpy(), hb, hr, hf,
                              host(ex.children))
        except Exception as e:  # pragma: no cover - reported below
            errors.append((k, repr(e)))

    threads = [threading.Thread(target=worker, args=(k,)) for k in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for k in range(4):
        gb, gr, gf, hb, hr, hf, ch = results[k]
        wb, wr, wf = want[k]
        np.testing.assert_array_equal(gb, wb)
        np.testing.assert_array_equal(gr, wr)
        np.testing.assert_array_equal(gf, wf)
        np.testing.assert_array_equal(hb, wb)
        np.testing.assert_array_equal(hr, wr)
        np.testing.assert_array_equal(hf, wf)
        np.testing.assert_array_equal(ch, want_x[k]['children'])


def test_shutdown_and_lazy_reinit(T, ops, oracle):
    b = synthetic_packed(10000, seed=3)
    a = np.random.default_rng(1).integers(0, 4, b.size).astype(np.uint8)
    want = oracle.step(b, a, None, seed=5)
    for _ in range(2):
        out = ops.step(dev(T, b), T.from_numpy(a).cuda(), None, seed=5)
        np.testing.assert_array_equal(host(out.boards), want[0])
        hb, _, _ = ops.step_host(b, a, None, seed=5)
        np.testing.assert_array_equal(hb, want[0])
        T.cuda.synchronize()
        ops.shutdown()
    ops.init()


def test_step_in_place_same_pointer(T, ops, oracle):
    """boards_out == boards_in (the documented aliasing): the LUT kernel and the general kernel."""
    for n in ((1 << 18) + 128, 3001):
        b = synthetic_packed(n, seed=n)
        a = np.random.default_rng(2).integers(0, 4, n).astype(np.uint8)
        ex = oracle.expand(b, child_layout=0)
        after = ex['after'].reshape(n, 4)[np.arange(n), a]
        s = valid_spawn_bytes(after, seed=4)
        want = oracle.step(b, a, s)
        d = dev(T, b)
        out = ops.StepOut(d, T.empty(n, dtype=T.int32, device='cuda'), T.empty(n, dtype=T.uint8, device='cuda'))
        ops.step(d, T.from_numpy(a).cuda(), T.from_numpy(s).cuda(), out=out)
        np.testing.assert_array_equal(host(d), want[0])
        np.testing.assert_array_equal(out.rewards.cpu().numpy(), want[1])


def test_pinned_blocks_and_copy_ceiling_leg(T, ops, oracle):
    n = (1 << 18) + 777
    hb = ops.pinned_empty(n, T.int64)
    ha, hs = ops.pinned_empty(n, T.uint8), ops.pinned_empty(n, T.uint8)
    hb[:] = synthetic_packed(n, seed=8).view(np.int64)
    ha[:] = np.random.default_rng(3).integers(0, 4, n)
    hs[:] = np.random.default_rng(4).integers(0, 32, n)
    out = (ops.pinned_empty(n, T.int64).view(np.uint64), ops.pinned_empty(n, T.int32), ops.pinned_empty(n, T.uint8))
    ops.step_host(hb, ha, hs, out=out)
    want = oracle.step(hb.view(np.uint64), ha, hs)
    np.testing.assert_array_equal(out[0], want[0])
    np.testing.assert_array_equal(out[2], want[2])
    ops.step_host_copies(hb, ha, hs, out)      # copies only: runs, leaves no error behind
    gpu_node, mem_node = ops.host_numa_info(hb)
    assert gpu_node >= -1 and mem_node >= -1
    del hb, ha, hs, out                        # finalizers return the blocks (g2048_host_free)
    import gc
    gc.collect()
    import ctypes
    assert ops.lib().g2048_host_free(ctypes.c_void_p(12345)) == -1   # not one of ours: EINVAL, no crash


def test_environment_wrapper_and_initial_board(T, oracle):
    """reinforce/game/env.py:248 Environment and core.py:324 create_initial_board."""
    from simulate_2048_b200 import game
    from simulate_2048_b200.game import core
    b = core.create_initial_board(123)
    assert tuple(b.shape) == (4, 4) and b.dtype == T.int32
    v = b.cpu().numpy()
    assert (v != 0).sum() == 2 and set(np.unique(v[v != 0])) <= {2, 4}
    want = oracle.unpack(oracle.reset(1, seed=123))[0]
    np.testing.assert_array_equal(v, want)
    many = core.create_initial_board(7, n=1000).cpu().numpy()
    np.testing.assert_array_equal(many, oracle.unpack(oracle.reset(1000, seed=7)))
    sp = T.tensor([3, 0x10 | 14], dtype=T.uint8, device='cuda')    # injected: rank 3 -> 2, rank 14 of 15 -> 4
    inj = core.create_initial_board(spawns=sp).cpu().numpy().reshape(-1)
    assert inj[3] == 2 and inj[15] == 4 and (inj != 0).sum() == 2

    env = game.Environment(seed=9)
    first = env.board.cpu().numpy().copy()
    assert (first != 0).sum() == 2 and not env.done
    assert tuple(env.observation().shape) == (16,) and tuple(env.legal_actions().shape) == (4,)
    total, steps = 0.0, 0
    rng = np.random.default_rng(0)
    while not env.done and steps < 2000:
        legal = np.flatnonzero(env.legal_actions().cpu().numpy())
        state, reward, done, info = env.step_env(int(legal[rng.integers(len(legal))]))
        total += reward
        steps += 1
        assert int(info['step_count']) == steps and float(info['total_reward']) == total
        assert int(info['max_tile']) == int(env.board.max())
    assert env.done and steps > 20 and not env.legal_actions().any()
    env2 = game.Environment(seed=9)          # same seed, same first board
    np.testing.assert_array_equal(env2.board.cpu().numpy(), first)
    env2.reset_env()
    assert int(env2.state.step_count) == 0 and not env2.done


def test_empty_batches_everywhere(T, ops):
    """n = 0 through every tensor-level entry point: no launch, no error, empty outputs."""
    from simulate_2048_b200 import _cabi
    e64 = T.empty(0, dtype=T.int64, device='cuda')
    e8 = T.empty(0, dtype=T.uint8, device='cuda')
    out = ops.step(e64, e8, None, seed=1)
    assert out.boards.numel() == 0 and out.rewards.numel() == 0 and out.flags.numel() == 0
    assert ops.step(e64, e8, e8).boards.numel() == 0
    assert ops.legal_flags(e64).numel() == 0 and tuple(ops.legal_actions(e64).shape) == (0, 4)
    for layout in ('none', 'dense', 'compact'):
        ex = ops.expand(e64, children=layout)
        assert ex.after.shape[0] == 0 and (ex.children is None or ex.children.numel() == 0)
    assert ops.expand(e64, children='compact').total.tolist() == [0, 0]
    ch, ne, ec = ops.after_state(e64)
    assert ch.numel() == 0 and ne.numel() == 0
    assert ops.reset(0, seed=1).numel() == 0 and ops.spawn(e64, None, seed=1).numel() == 0
    obs, mx = ops.observe(e64)
    assert tuple(obs.shape) == (0, 16) and ops.observe_onehot(e64).numel() == 0
    assert ops.unpack(e64).numel() == 0 and ops.pack(T.empty((0, 4, 4), dtype=T.int32, device='cuda')).numel() == 0
    r = ops.rollout(0, seed=1)
    assert r.score.numel() == 0 and int(r.stats.sum()) == 0
    ops.env_step(e64, T.empty(0, dtype=T.int32, device='cuda'), e8, T.empty(0, dtype=T.float32, device='cuda'), e8, None, seed=1)
    ops.selfplay_step(e64, T.empty(0, dtype=T.int32, device='cuda'), e8, T.empty(0, dtype=T.float32, device='cuda'),
                      T.empty((0, 4), dtype=T.float32, device='cuda'), seed=1)
    assert ops.sample_action(T.empty((0, 4), dtype=T.float32, device='cuda'), T.empty((0, 4), dtype=T.bool, device='cuda')).numel() == 0
    hb, hr, hf = ops.step_host(np.empty(0, np.uint64), np.empty(0, np.uint8), None, seed=1)
    assert hb.size == 0 and hr.size == 0 and hf.size == 0
    # bad arguments come back as error codes, not crashes
    lib = ops.lib()
    assert lib.g2048_step(None, None, None, 0, 0, None, 0, None, None, None, None, None, 16, None) == _cabi.EINVAL
    assert lib.g2048_expand_compact(None, None, None, None, None, None, None, None, 0, None, 4, None) == _cabi.EINVAL
    assert lib.g2048_observe_onehot(e64.data_ptr(), e8.data_ptr(), 7, 1, None) == _cabi.EINVAL
