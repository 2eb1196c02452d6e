"""GPU parity of the MCTS-leaf expansion (config 3) and the fused rollout (config 4)."""

import numpy as np
import pytest

from tests.helpers import pack_tiles, synthetic_packed

pytestmark = pytest.mark.gpu

FLAG_DONE, FLAG_MOVED, FLAG_OVERFLOW = 0x10, 0x20, 0x40


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
    a = np.ascontiguousarray(a)
    if a.dtype == np.uint64:
        a = a.view(np.int64)
    return T.from_numpy(a).cuda()


def host(t):
    a = t.cpu().numpy()
    return a.view(np.uint64) if a.dtype == np.int64 else a


# ------------------------------------------------------------------ expand

def test_expand_reproduces_reference(T, ops, golden):
    """mask + latent_state x4 + after_state children of the real reference (fixture)."""
    g = golden('synthetic')
    ex = ops.expand(dev(T, g['packed']), children='compact')
    np.testing.assert_array_equal(host(ex.after), g['after'])
    np.testing.assert_array_equal(ex.rewards.cpu().numpy(), g['rewards'].astype(np.int32))
    np.testing.assert_array_equal(ex.mask.cpu().numpy() & 0xF, g['mask'])
    np.testing.assert_array_equal(ex.child_count.cpu().numpy(), g['child_count'])
    np.testing.assert_array_equal(host(ex.children), g['children'])
    assert int(ex.child_count.sum().item()) == 249022


@pytest.mark.parametrize('n', [1, 31, 33, 65536, 65536 + 7, (1 << 17) + 37])
def test_expand_vs_oracle(T, ops, oracle, n):
    boards = synthetic_packed(n, seed=n + 1, max_exp=6 if n % 2 else 11)
    for layout, code in (('none', 0), ('dense', 1), ('compact', 2)):
        ex = ops.expand(dev(T, boards), children=layout)
        want = oracle.expand(boards, child_layout=code)
        np.testing.assert_array_equal(host(ex.after), want['after'])
        np.testing.assert_array_equal(ex.rewards.cpu().numpy(), want['rewards'])
        np.testing.assert_array_equal(ex.mask.cpu().numpy(), want['mask'])
        np.testing.assert_array_equal(ex.num_empty.cpu().numpy(), want['num_empty'])
        np.testing.assert_array_equal(ex.child_count.cpu().numpy(), want['child_count'])
        if code == 1:
            np.testing.assert_array_equal(host(ex.children), want['children'])
        if code == 2:
            np.testing.assert_array_equal(ex.child_offsets.cpu().numpy(), want['child_offsets'])
            np.testing.assert_array_equal(host(ex.children), want['children'])


def test_expand_edge_cases(T, ops, oracle):
    tiles = np.array([
        [[0] * 4] * 4,
        [[2, 4, 2, 4], [4, 2, 4, 2], [2, 4, 2, 4], [4, 2, 4, 2]],
        [[2, 2, 4, 8], [16, 32, 64, 128], [256, 512, 1024, 2048], [4096, 8192, 16384, 32768]],
        [[32768, 32768, 0, 0], [32768, 0, 0, 0], [0] * 4, [0] * 4],
        [[2, 4, 8, 16], [32, 64, 128, 0], [256, 512, 1024, 2048], [4096, 8192, 16384, 32768]],
    ])
    boards = pack_tiles(tiles)
    ex = ops.expand(dev(T, boards), children='dense')
    want = oracle.expand(boards, child_layout=1)
    np.testing.assert_array_equal(host(ex.after), want['after'])
    np.testing.assert_array_equal(ex.mask.cpu().numpy(), want['mask'])
    np.testing.assert_array_equal(ex.child_count.cpu().numpy(), want['child_count'])
    np.testing.assert_array_equal(host(ex.children), want['children'])
    m = ex.mask.cpu().numpy()
    assert m[0] == FLAG_DONE and m[1] == FLAG_DONE and m[3] & FLAG_OVERFLOW
    # a legal move onto a full afterstate cannot exist (a move frees or keeps an empty cell);
    # the single-empty board has 2 children per legal action with P = 0.9 / 0.1
    cc = ex.child_count.cpu().numpy()
    assert set(cc[4][cc[4] > 0]) == {2}


def test_expand_large_batch_with_big_tiles(T, ops, oracle):
    """n >= 2^17 runs the persistent moves kernel with the staged table; boards holding a tile
    >= 16384 (also 32768 pairs that overflow) take its exact path inside the same launch."""
    n = (1 << 17) + 5
    boards = synthetic_packed(n, seed=99, max_exp=11)
    boards[::97] = synthetic_packed(len(boards[::97]), seed=5, max_exp=15)
    ex = ops.expand(dev(T, boards), children='compact')
    want = oracle.expand(boards, child_layout=2)
    np.testing.assert_array_equal(host(ex.after), want['after'])
    np.testing.assert_array_equal(ex.rewards.cpu().numpy(), want['rewards'])
    np.testing.assert_array_equal(ex.mask.cpu().numpy(), want['mask'])
    np.testing.assert_array_equal(ex.child_count.cpu().numpy(), want['child_count'])
    np.testing.assert_array_equal(host(ex.children), want['children'])
    assert (ex.mask.cpu().numpy() & FLAG_OVERFLOW).any()


def test_expand_odd_child_offsets(T, ops):
    """Caller-supplied compact offsets that break the 16-byte alignment of the child pairs take
    the 64-bit store path: same children, shifted by one slot."""
    boards = dev(T, synthetic_packed(5000, seed=17))
    ex = ops.expand(boards, children='compact')
    total = ex.children.numel()
    shifted = ex._replace(children=T.full((total + 1,), -1, dtype=T.int64, device='cuda'),
                          child_offsets=ex.child_offsets + 1)
    ops.expand(boards, children='compact', out=shifted, child_offsets=ex.child_offsets + 1)
    assert int(shifted.children[0].item()) == -1
    np.testing.assert_array_equal(host(shifted.children[1:]), host(ex.children))


@pytest.mark.parametrize('n', [1, 127, 129, 40000, (1 << 17) + 1025, (1 << 19) + 33 * 1024 + 7])
def test_expand_compact_device_scan(T, ops, oracle, n):
    """g2048_expand_compact: offsets scanned on the device (single-launch look-back below 2^17
    boards; group sums from the moves kernel + tile prefixes above), no host round trip; buffers of an earlier call are reused
    with DIFFERENT boards (the offsets are recomputed, never stale); a capacity that is too small
    raises the overflow flag and never writes past the buffer."""
    b1 = synthetic_packed(n, seed=3 * n + 1)
    b2 = synthetic_packed(n, seed=3 * n + 2, max_exp=4)
    w1, w2 = oracle.expand(b1, child_layout=2), oracle.expand(b2, child_layout=2)
    cap = max(len(w1['children']), len(w2['children'])) + 5
    ex = ops.expand(dev(T, b1), children='compact', capacity=cap)
    assert ex.total.tolist() == [len(w1['children']), 0]
    np.testing.assert_array_equal(ex.child_offsets.cpu().numpy(), w1['child_offsets'])
    np.testing.assert_array_equal(host(ex.children[:len(w1['children'])]), w1['children'])
    ex.children.fill_(-1)
    ex2 = ops.expand(dev(T, b2), children='compact', out=ex)       # same buffers, other boards
    assert ex2.total.tolist() == [len(w2['children']), 0]
    np.testing.assert_array_equal(ex2.child_offsets.cpu().numpy(), w2['child_offsets'])
    np.testing.assert_array_equal(host(ex2.children[:len(w2['children'])]), w2['children'])
    np.testing.assert_array_equal(ex2.child_count.cpu().numpy(), w2['child_count'])
    assert (host(ex2.children[len(w2['children']):]).view(np.int64) == -1).all()
    # too small: boards that fit are written, nothing beyond the capacity is touched
    small = max(len(w1['children']) // 2, 1)
    guard = T.full((small + 64,), -1, dtype=T.int64, device='cuda')
    tight = ex._replace(children=guard[:small])
    got = ops.expand(dev(T, b1), children='compact', out=tight)
    assert got.total.tolist() == [len(w1['children']), 1]
    assert (guard[small:].cpu().numpy() == -1).all()
    fits = w1['child_offsets'] + w1['child_count'].reshape(n, 4).sum(axis=1) <= small
    k = n if fits.all() else int(np.argmin(fits))   # offsets are monotone: the first k boards fit
    if k:
        end = int(w1['child_offsets'][k - 1] + w1['child_count'].reshape(n, 4)[k - 1].sum())
        np.testing.assert_array_equal(host(guard[:end]), w1['children'][:end])


def test_expand_dense_writes_every_slot(T, ops, oracle):
    """The dense layout needs no memset: slots beyond child_count are written as 0 by the kernel,
    also for done boards, illegal actions and boards on the exact (big tile) path."""
    boards = synthetic_packed(3000, seed=23, max_exp=11)
    boards[::50] = synthetic_packed(len(boards[::50]), seed=6, max_exp=15)
    boards[:2] = pack_tiles(np.array([[[0] * 4] * 4, [[2, 4, 2, 4], [4, 2, 4, 2], [2, 4, 2, 4], [4, 2, 4, 2]]]))
    d = dev(T, boards)
    ex = ops.expand(d, children='dense')
    ex.children.fill_(-1)
    ops.expand(d, children='dense', out=ex)
    np.testing.assert_array_equal(host(ex.children), oracle.expand(boards, child_layout=1)['children'])


def test_expand_children_from_earlier_afterstates(T, ops, oracle):
    """g2048_expand_children rebuilds the children from the afterstates and counts of an earlier
    call (the moves are paid once): identical to the one-call result, dense and compact, and the
    argument checks of the entry point."""
    from simulate_2048_b200 import _cabi
    boards = synthetic_packed(7001, seed=31)
    boards[::97] = synthetic_packed(len(boards[::97]), seed=8, max_exp=15)   # exact-path boards
    d = dev(T, boards)
    n = d.numel()
    lib = ops.lib()
    ptr = lambda t: None if t is None else t.data_ptr()
    for layout, code in (('dense', 1), ('compact', 2)):
        want = oracle.expand(boards, child_layout=code)
        fixed = ops.expand(d, children='none')
        np.testing.assert_array_equal(host(fixed.after), want['after'].reshape(n, 4))
        if code == 1:
            ch, off = T.full((n, 4, 16, 2), -1, dtype=T.int64, device='cuda'), None
        else:
            per_env = fixed.child_count.to(T.int64).sum(dim=1)
            off = (T.cumsum(per_env, 0) - per_env).contiguous()
            ch = T.full((int(per_env.sum().item()),), -1, dtype=T.int64, device='cuda')
        rc = lib.g2048_expand_children(ptr(fixed.after), ptr(fixed.child_count), ptr(off), ptr(ch), code, n, None)
        assert rc == 0
        T.cuda.synchronize()
        np.testing.assert_array_equal(host(ch).reshape(-1), np.asarray(want['children']).reshape(-1))
    assert lib.g2048_expand_children(None, ptr(fixed.child_count), ptr(off), ptr(ch), 2, n, None) == _cabi.EINVAL
    assert lib.g2048_expand_children(ptr(fixed.after), ptr(fixed.child_count), None, ptr(ch), 2, n, None) == _cabi.EINVAL
    assert lib.g2048_expand_children(ptr(fixed.after), ptr(fixed.child_count), ptr(off), ptr(ch), 0, n, None) == _cabi.EINVAL
    assert lib.g2048_expand_children(ptr(fixed.after), ptr(fixed.child_count), ptr(off), ptr(ch), 2, 0, None) == 0


def test_after_state_vs_oracle(T, ops, oracle):
    boards = synthetic_packed(20000, seed=3)
    boards[:3] = pack_tiles(np.array([[[0] * 4] * 4, [[2] * 4] * 4, [[2, 4, 8, 16], [32, 64, 128, 0], [2, 4, 8, 16], [32, 64, 128, 256]]]))
    ch, ne, ec = ops.after_state(dev(T, boards))
    o_ch, o_ne, o_ec = oracle.after_state(boards)
    np.testing.assert_array_equal(host(ch).reshape(-1, 32), o_ch)
    np.testing.assert_array_equal(ne.cpu().numpy(), o_ne)
    np.testing.assert_array_equal(ec.cpu().numpy().view(np.uint16), o_ec)
    # every slot is written (no memset needed), also for ragged sizes and a children pointer that is
    # only 8-byte aligned (scalar kernel); nothing outside the n * 32 slots is touched
    lib = ops.lib()
    for n in (1, 17, 33, 4001):
        d = dev(T, boards[:n])
        for off in (0, 1):
            buf = T.full((n * 32 + 4,), -1, dtype=T.int64, device='cuda')
            view = buf[2 + off:2 + off + n * 32]
            assert lib.g2048_after_state(d.data_ptr(), view.data_ptr(), None, None, n, None) == 0
            T.cuda.synchronize()
            np.testing.assert_array_equal(host(view).reshape(n, 32), o_ch[:n])
            assert int(buf[1 + off].item()) == -1 and int(buf[2 + off + n * 32].item()) == -1


# ------------------------------------------------------------------ rollout

@pytest.mark.parametrize('n', [1, 100, 20000, 148 * 512 + 13])
def test_rollout_vs_oracle(T, ops, oracle, n):
    """Every episode is bit-identical to the oracle's (same Philox stream keyed by env, step):
    final board, score, length, max tile and the statistics vector."""
    kw = dict(seed=0xABCDEF, env_offset=10**12)
    out = ops.rollout(n, **kw)
    want = oracle.rollout(n, **kw)
    np.testing.assert_array_equal(host(out.boards), want['boards'])
    np.testing.assert_array_equal(out.score.cpu().numpy(), want['score'])
    np.testing.assert_array_equal(out.steps.cpu().numpy(), want['steps'])
    np.testing.assert_array_equal(out.maxexp.cpu().numpy(), want['maxexp'])
    np.testing.assert_array_equal(out.stats.cpu().numpy().view(np.uint64), want['stats'])
    assert ((ops.legal_flags(out.boards).cpu().numpy() & FLAG_DONE) != 0).all()


def test_rollout_trace_replays_through_step(T, ops, oracle):
    """Trace mode: (action, spawn cell, value) per step replayed through g2048_step with
    injected absolute-cell spawns AND through the oracle reproduces the episode."""
    n, stride = 256, 4096
    out = ops.rollout(n, seed=7, trace_stride=stride)
    trace = out.trace.cpu().numpy()
    want = oracle.rollout(n, seed=7, trace_stride=stride)
    np.testing.assert_array_equal(trace, want['trace'])
    boards = oracle.reset(n, seed=7)
    score = np.zeros(n, np.int64)
    alive = np.ones(n, bool)
    for t in range(int(out.steps.max().item())):
        byte = trace[:, t]
        alive &= byte != 0
        actions = (byte & 3).astype(np.uint8)
        spawns = (0x80 | ((byte >> 2) & 0xF) | np.where(byte & 0x40, 0x10, 0)).astype(np.uint8)
        res = ops.step(dev(T, boards), dev(T, actions), dev(T, spawns), done_in=dev(T, (~alive).astype(np.uint8)))
        boards = host(res.boards)
        score += np.where(alive, res.rewards.cpu().numpy(), 0)
    np.testing.assert_array_equal(boards, host(out.boards))
    np.testing.assert_array_equal(score, out.score.cpu().numpy())
    # the start boards are g2048_reset's (words 0, 1 of each env's game stream)
    np.testing.assert_array_equal(host(ops.reset(n, seed=7)), oracle.reset(n, seed=7))


def test_rollout_sharding_invariance_and_caps(T, ops, oracle):
    n = 6000
    whole = ops.rollout(n, seed=3)
    a = ops.rollout(2500, seed=3, env_offset=0)
    b = ops.rollout(3500, seed=3, env_offset=2500)
    np.testing.assert_array_equal(T.cat([a.score, b.score]).cpu().numpy(), whole.score.cpu().numpy())
    from simulate_2048_b200.parallel import combine_stats
    np.testing.assert_array_equal(combine_stats(T.stack([a.stats, b.stats])).cpu().numpy(),
                                  whole.stats.cpu().numpy())
    capped = ops.rollout(1000, seed=3, max_steps=50)
    want = oracle.rollout(1000, seed=3, max_steps=50)
    np.testing.assert_array_equal(capped.steps.cpu().numpy(), want['steps'])
    np.testing.assert_array_equal(capped.stats.cpu().numpy().view(np.uint64), want['stats'])
    assert int(capped.stats[5].item()) > 0
    # from given boards
    start = synthetic_packed(3000, seed=12)
    got = ops.rollout(3000, seed=1, boards=dev(T, start))
    want = oracle.rollout(3000, seed=1, from_reset=False, boards=start)
    np.testing.assert_array_equal(host(got.boards), want['boards'])
    np.testing.assert_array_equal(got.score.cpu().numpy(), want['score'])


def test_rollout_stats_match_oracle_fixture(T, ops):
    """The statistics vector of BASELINE config 4 at 2^16, 2^20 and 2^21 envs (seed 2048) equals the
    oracle-computed vector bench.py checks the sharded run against — also when the range is split
    in two unequal shards and combined as parallel.reduce_stats does."""
    import json
    import os
    from simulate_2048_b200.parallel import combine_stats
    here = os.path.dirname(os.path.abspath(__file__))
    fix = json.load(open(os.path.join(here, 'golden', 'rollout_stats.json')))
    for n in (1 << 16, 1 << 20, 1 << 21):
        out = ops.rollout(n, seed=fix['seed'], per_env=False)
        assert [int(x) for x in out.stats.cpu().numpy().view(np.uint64)] == fix['stats'][str(n)]
    n, cut = 1 << 20, 300001
    a = ops.rollout(cut, seed=fix['seed'], per_env=False)
    b = ops.rollout(n - cut, seed=fix['seed'], env_offset=cut, per_env=False)
    both = combine_stats(T.stack([a.stats, b.stats]))
    assert [int(x) for x in both.cpu().numpy().view(np.uint64)] == fix['stats'][str(n)]


def test_rollout_statistics_sanity(T, ops):
    """Random play statistics agree with the reference engine's (mean length
    ~117, mean score ~1.1k; max tile mostly 64..256)."""
    from simulate_2048_b200.parallel import EpisodeStats
    out = ops.rollout(200000, seed=11, per_env=False)
    s = EpisodeStats.from_vector(out.stats.cpu()).summary()
    assert 105 < s['mean_length'] < 130 and 950 < s['mean_reward'] < 1250
    assert s['max_tile'] <= 2048 and s['reached_2048'] < 20


def test_rollout_persistent_kernel_variants(T, ops, oracle):
    """The shared-memory-table rollout kernel (n >= 148 * 512 envs) in its other template variants:
    games that start from given boards (incl. boards that are already over and a step cap), and
    trace mode with a stride shorter than the games (truncated traces keep their terminator)."""
    n = 148 * 512 + 77
    start = synthetic_packed(n, seed=21, max_exp=6)
    start[::1000] = 0x1212212112122121                      # no legal move: emitted at once
    got = ops.rollout(n, seed=9, env_offset=5, boards=dev(T, start), max_steps=300)
    want = oracle.rollout(n, seed=9, env_offset=5, from_reset=False, boards=start, max_steps=300)
    np.testing.assert_array_equal(host(got.boards), want['boards'])
    np.testing.assert_array_equal(got.score.cpu().numpy(), want['score'])
    np.testing.assert_array_equal(got.steps.cpu().numpy(), want['steps'])
    np.testing.assert_array_equal(got.maxexp.cpu().numpy(), want['maxexp'])
    np.testing.assert_array_equal(got.stats.cpu().numpy().view(np.uint64), want['stats'])
    assert (got.steps.cpu().numpy()[::1000] == 0).all()
    stride = 48
    tr = ops.rollout(n, seed=4, trace_stride=stride, max_steps=200)
    wt = oracle.rollout(n, seed=4, trace_stride=stride, max_steps=200)
    np.testing.assert_array_equal(tr.trace.cpu().numpy(), wt['trace'])
    np.testing.assert_array_equal(tr.steps.cpu().numpy(), wt['steps'])
    np.testing.assert_array_equal(tr.stats.cpu().numpy().view(np.uint64), wt['stats'])


@pytest.mark.gpu
@pytest.mark.parametrize('threads', [512, 768])
def test_rollout_other_cta_sizes(threads):
    """The 512- and 768-thread instantiations of the persistent rollout kernel (the library picks
    640; G2048_ROLL_THREADS selects the others for measurements) give the oracle's episodes too.
    The knob is read when the library loads, so the check runs in a child process."""
    import os
    import subprocess
    import sys
    code = (
        'import sys; sys.path.insert(0, ".")\n'
        'import numpy as np, torch\n'
        'from oracle import c_oracle\n'
        'from simulate_2048_b200 import ops\n'
        'n = 148 * 768 + 13\n'
        'out = ops.rollout(n, seed=77, env_offset=3)\n'
        'want = c_oracle.rollout(n, seed=77, env_offset=3)\n'
        'assert (out.boards.cpu().numpy().view(np.uint64) == want["boards"]).all()\n'
        'assert (out.score.cpu().numpy() == want["score"]).all() and (out.steps.cpu().numpy() == want["steps"]).all()\n'
        'assert (out.stats.cpu().numpy().view(np.uint64) == want["stats"]).all()\n'
        'print("ok")\n')
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, G2048_ROLL_THREADS=str(threads))
    r = subprocess.run([sys.executable, '-c', code], cwd=root, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.strip().endswith('ok'), r.stderr[-2000:]


@pytest.mark.parametrize('n', [3000, 148 * 640 + 41])
def test_rollout_big_tiles_leave_and_rejoin_the_fast_path(T, ops, oracle, n):
    """Rollouts from boards that hold, or are about to build, tiles >= 16384: such a board's rows lie
    outside the staged table window, so its moves take the exact out-of-line path.  The kernel does
    not test the board every step: a game carries the OR of its move scores (>= 16384 iff a tile
    >= 16384 may have been built) and the out-of-line path re-evaluates the board exactly.  Covered:
    boards that start with 16384 / 32768 tiles (incl. 32768 pairs that overflow), boards of 4096 /
    8192 tiles that cross the threshold mid-game, both kernel variants (table gathered from L2 /
    staged in shared memory)."""
    rng = np.random.default_rng(n)
    start = synthetic_packed(n, seed=n, max_exp=15)                       # anything up to 32768
    dense = rng.integers(12, 14, (n // 3, 16)).astype(np.uint64)          # full boards of 4096 / 8192: merge chains
    dense[rng.random(dense.shape) < 0.15] = 0
    start[: n // 3] = np.bitwise_or.reduce(dense << (4 * np.arange(16, dtype=np.uint64))[None, :], axis=1)
    pairs = rng.integers(14, 16, (n // 8, 16)).astype(np.uint64)          # 16384 / 32768 everywhere: overflows
    start[n // 3: n // 3 + n // 8] = np.bitwise_or.reduce(pairs << (4 * np.arange(16, dtype=np.uint64))[None, :], axis=1)
    got = ops.rollout(n, seed=31, env_offset=9, boards=dev(T, start), max_steps=500)
    want = oracle.rollout(n, seed=31, env_offset=9, from_reset=False, boards=start, max_steps=500)
    np.testing.assert_array_equal(host(got.boards), want['boards'])
    np.testing.assert_array_equal(got.score.cpu().numpy(), want['score'])
    np.testing.assert_array_equal(got.steps.cpu().numpy(), want['steps'])
    np.testing.assert_array_equal(got.maxexp.cpu().numpy(), want['maxexp'])
    np.testing.assert_array_equal(got.stats.cpu().numpy().view(np.uint64), want['stats'])
    st = want['stats']
    assert st[6] > 0 and st[8 + 15] > 0          # some games overflowed, some ended holding a 32768
    # the threshold is crossed mid-game: games that started below 16384 and ended with one
    began_small = oracle.unpack(start).reshape(n, -1).max(axis=1) < 16384
    assert (began_small & (want['maxexp'] >= 14)).sum() > 0
