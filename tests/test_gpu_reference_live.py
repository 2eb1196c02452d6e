"""The CUDA path against the REAL reference engine (oracle/_ref, a verbatim copy of the reference's
twentyfortyeight/ package staged by oracle/make_ref.py) on fresh random boards — no fixtures, no
restatement in between.  Through the C ABI (ops.* are one-line ctypes calls)."""

import numpy as np
import pytest

from tests.helpers import pack_tiles, synthetic_packed

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def ref():
    from oracle import make_ref
    if not make_ref.available():
        pytest.skip('oracle/_ref not staged')
    return make_ref.load()


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


@pytest.mark.parametrize('seed,max_exp', [(101, 11), (202, 4), (303, 15)])
def test_step_mask_done_equal_reference(T, ops, oracle, ref, seed, max_exp):
    """next_state (gameboard.py:291) with the reference's own generator: the spawned tile is read off
    the reference's result and injected; boards, rewards, legal mask (gamemove.py:45) and is_done
    (gameboard.py:328) of the new board must be equal."""
    gb, gm, _ = ref
    n = 600
    boards = synthetic_packed(n, seed=seed, max_exp=max_exp)
    tiles = oracle.unpack(boards).astype(np.int64).reshape(n, 4, 4)
    rng = np.random.default_rng(seed)
    actions = rng.integers(0, 4, n).astype(np.uint8)
    gb._GENERATOR = np.random.default_rng(seed + 1)
    spawns = np.zeros(n, np.uint8)
    want_tiles = np.zeros_like(tiles)
    want_reward = np.zeros(n, np.int64)
    keep = np.ones(n, bool)
    for i in range(n):
        after, _ = gb.latent_state(tiles[i], int(actions[i]))
        if np.asarray(after).max() > 32768:      # a 32768+32768 merge: outside the packed format
            keep[i] = False
            continue
        nxt, r = gb.next_state(tiles[i].copy(), int(actions[i]))
        diff = np.argwhere(np.asarray(nxt) != np.asarray(after))
        assert len(diff) <= 1
        if len(diff):
            rr, cc = int(diff[0][0]), int(diff[0][1])
            spawns[i] = 0x80 | (4 * rr + cc) | (0x10 if nxt[rr, cc] == 4 else 0)
        want_tiles[i], want_reward[i] = nxt, int(r)
    out = ops.step(dev(T, boards), T.from_numpy(actions).cuda(), T.from_numpy(spawns).cuda())
    got_tiles = ops.unpack(out.boards).cpu().numpy()
    flags = out.flags.cpu().numpy()
    np.testing.assert_array_equal(got_tiles[keep], want_tiles[keep])
    np.testing.assert_array_equal(out.rewards.cpu().numpy()[keep], want_reward[keep])
    for i in np.flatnonzero(keep):
        mask = gm.legal_actions_mask(want_tiles[i])
        assert tuple(bool(flags[i] & (1 << a)) for a in range(4)) == tuple(bool(m) for m in mask)
        if want_tiles[i].any():                  # the empty board is the one divergence (JAX: done)
            assert bool(flags[i] & 0x10) == bool(gb.is_done(want_tiles[i]))


def test_expand_equals_reference(T, ops, oracle, ref):
    """latent_state x4 (gameboard.py:105) and after_state (gameboard.py:132) of every legal
    afterstate: afterstates, rewards, child order and probabilities."""
    gb, gm, _ = ref
    n = 300
    boards = synthetic_packed(n, seed=77, max_exp=9)
    tiles = oracle.unpack(boards).astype(np.int64).reshape(n, 4, 4)
    ex = ops.expand(dev(T, boards), children='compact')
    after = ops.unpack(ex.after.reshape(-1)).cpu().numpy().reshape(n, 4, 4, 4)
    rewards = ex.rewards.cpu().numpy()
    children = ops.unpack(ex.children).cpu().numpy()
    counts, offs = ex.child_count.cpu().numpy(), ex.child_offsets.cpu().numpy()
    num_empty = ex.num_empty.cpu().numpy()
    for i in range(n):
        mask = gm.legal_actions_mask(tiles[i])
        k = int(offs[i])
        for a in range(4):
            s, r = gb.latent_state(tiles[i], a)
            np.testing.assert_array_equal(after[i, a], s)
            assert rewards[i, a] == int(r)
            if not mask[a]:
                assert counts[i, a] == 0
                continue
            kids = gb.after_state(np.array(s))
            assert counts[i, a] == len(kids)
            for j, (child, p) in enumerate(kids):
                np.testing.assert_array_equal(children[k + j], child)
                assert p == pytest.approx((0.9 if j % 2 == 0 else 0.1) / num_empty[i, a])
            k += len(kids)


def test_game_through_mirror_equals_reference_env(T, ref):
    """TwentyFortyEight.step (envs/twentyfortyeight.py:104) of the reference against the device-backed
    mirror class over whole games: same observation, reward and termination when the mirror is fed
    the reference's spawns."""
    gb, gm, RefEnv = ref
    from simulate_2048_b200.twentyfortyeight.core import gameboard as mirror
    from simulate_2048_b200.twentyfortyeight.core import gamemove as mirror_m
    rng = np.random.default_rng(5)
    for g in range(3):
        gb._GENERATOR = np.random.default_rng(g)
        env = RefEnv()
        state = np.array(env.reset(seed=g))
        steps = 0
        while not env.is_finished and steps < 400:
            legal = gm.legal_actions(state)
            a = int(legal[rng.integers(len(legal))])
            after_ref, r_ref = gb.latent_state(state, a)
            after_dev, r_dev = mirror.latent_state(state.copy(), a)
            np.testing.assert_array_equal(after_dev, after_ref)
            assert r_dev == r_ref
            assert tuple(mirror_m.legal_actions_mask(state)) == tuple(gm.legal_actions_mask(state))
            obs, _, done = env.step(a)
            state = np.array(obs)
            assert mirror.is_done(state) == done
            steps += 1
        assert env.is_finished


def test_rollout_traces_replay_through_the_reference_engine(T, ops, oracle, ref):
    """SURVEY.md §8d config 4 parity: the fused rollout kernel emits (action, spawn cell, tile) per
    step; replayed through the REAL reference engine — ``latent_state`` for the move, the spawned
    tile written where the trace says, ``legal_actions`` to confirm the action was legal and
    ``is_done`` at the end — every game ends on the kernel's final board with the kernel's score."""
    gb, gm, _ = ref
    n, stride = 96, 2048
    out = ops.rollout(n, seed=2024, env_offset=17, trace_stride=stride)
    trace = out.trace.cpu().numpy()
    start = ops.unpack(ops.reset(n, seed=2024, env_offset=17)).cpu().numpy().astype(np.int64)
    final = ops.unpack(out.boards).cpu().numpy()
    scores, steps = out.score.cpu().numpy(), out.steps.cpu().numpy()
    for i in range(n):
        board, score, t = start[i].copy(), 0, 0
        while trace[i, t]:
            byte = int(trace[i, t])
            action, cell, four = byte & 3, (byte >> 2) & 0xF, (byte >> 6) & 1
            assert action in gm.legal_actions(board)
            after, reward = gb.latent_state(board, action)
            board = np.array(after)
            assert board[cell // 4, cell % 4] == 0
            board[cell // 4, cell % 4] = 4 if four else 2
            score += int(reward)
            t += 1
        assert t == steps[i] and score == scores[i]
        np.testing.assert_array_equal(board, final[i])
        assert gb.is_done(board) and not gm.legal_actions(board)
