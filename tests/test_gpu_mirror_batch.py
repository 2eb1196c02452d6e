"""Batched entry points of the NumPy-engine mirror (twentyfortyeight/core/batch.py) and the lock-step
vector env (envs/vector.py) against the single-board semantics of the reference engine
(oracle/_ref when staged, else the NumPy restatement) and the C oracle."""

import numpy as np
import pytest

from tests.helpers import synthetic_packed, valid_spawn_bytes

pytestmark = pytest.mark.gpu


def _engine():
    from oracle import make_ref
    if make_ref.available():
        gb, gm, _ = make_ref.load()
        return gb, gm
    from oracle import np_engine
    return np_engine, np_engine


def test_batch_functions_equal_single_board_reference(oracle):
    from simulate_2048_b200.twentyfortyeight.core import batch
    gb, gm = _engine()
    n = 500
    packed = synthetic_packed(n, seed=61, max_exp=9)
    tiles = oracle.unpack(packed).astype(np.int64).reshape(n, 4, 4)
    actions = np.random.default_rng(0).integers(0, 4, n)
    after, rew = batch.latent_state_batch(tiles, actions)
    masks = batch.legal_actions_mask_batch(tiles)
    done = batch.is_done_batch(tiles)
    kids, probs, counts = batch.after_state_batch(tiles)
    assert after.dtype == np.int64 and after.shape == (n, 4, 4) and kids.shape == (n, 32, 4, 4)
    for i in range(n):
        s, r = gb.latent_state(tiles[i], int(actions[i]))
        np.testing.assert_array_equal(after[i], s)
        assert rew[i] == r
        assert tuple(masks[i]) == tuple(bool(x) for x in gm.legal_actions_mask(tiles[i]))
        assert bool(done[i]) == bool(gb.is_done(tiles[i]))
        want = gb.after_state(tiles[i])
        assert counts[i] == len(want)
        for j, (child, p) in enumerate(want):
            np.testing.assert_array_equal(kids[i, j], child)
            assert probs[i, j] == pytest.approx(p)
        assert not kids[i, len(want):].any() and not probs[i, len(want):].any()
    # full transition with injected spawns: the oracle's boards, rewards and moved flags
    ex = oracle.expand(packed, child_layout=0)
    spawns = valid_spawn_bytes(ex['after'].reshape(n, 4)[np.arange(n), actions], seed=2)
    new, r2, moved = batch.next_state_batch(tiles, actions, spawns=spawns)
    wb, wr, wf = oracle.step(packed, actions.astype(np.uint8), spawns)
    np.testing.assert_array_equal(new, oracle.unpack(wb).reshape(n, 4, 4))
    np.testing.assert_array_equal(r2, wr)
    np.testing.assert_array_equal(moved, (wf & 0x20) != 0)
    same = batch.next_state_batch(tiles, actions, seed=7)[0]
    np.testing.assert_array_equal(same, batch.next_state_batch(tiles, actions, seed=7)[0])   # seeded: repeatable
    empty = np.zeros((3, 4, 4), dtype=np.int64)
    assert not batch.is_done_batch(empty).any()               # NumPy engine: an empty board is not done


def test_vector_env_plays_config1_games(oracle):
    """1024 random-policy games to game-over through VectorTwentyFortyEight (BASELINE configs[0]):
    every transition replays through the oracle (same Philox key), rewards and termination agree,
    and the statistics are those of random 2048 play."""
    from simulate_2048_b200.twentyfortyeight.envs import VectorTwentyFortyEight
    n = 1024
    env = VectorTwentyFortyEight(n, seed=5)
    obs = env.reset(seed=5)
    assert obs.shape == (n, 4, 4) and obs.dtype == np.int64 and ((obs != 0).sum(axis=(1, 2)) == 2).all()
    key = env._key
    st = (oracle.reset(n, seed=key), np.zeros(n, np.int32), np.zeros(n, np.uint8), np.zeros(n, np.float32))
    np.testing.assert_array_equal(obs.reshape(n, 16), oracle.unpack(st[0]).reshape(n, 16))
    rng = np.random.default_rng(0)
    steps = np.zeros(n, np.int64)
    total = np.zeros(n, np.float64)
    it = 0
    while not env.is_finished.all() and it < 3000:
        mask = env.legal_actions_mask
        pick = rng.random((n, 4)) * mask
        actions = pick.argmax(axis=1).astype(np.uint8)            # uniform over the legal actions
        alive = ~env.is_finished
        obs, rew, fin = env.step(actions)
        st, o_rew, o_fl, _, _ = oracle.env_step(*st, actions, None, seed=key)
        if it % 20 == 0 or fin.all():
            np.testing.assert_array_equal(obs.reshape(n, 16), oracle.unpack(st[0]).reshape(n, 16))
            np.testing.assert_array_equal(fin, st[2].astype(bool))
        np.testing.assert_array_equal(rew, o_rew)
        steps += alive
        total += rew
        it += 1
    assert env.is_finished.all()
    np.testing.assert_array_equal(steps, st[1])
    np.testing.assert_allclose(env.total_reward, total)
    assert 105 < steps.mean() < 130 and 950 < total.mean() < 1250      # ~118 moves, ~1.1k points
    enc = VectorTwentyFortyEight(8, encoded=True, normalize=True, seed=1)
    o, r, _ = enc.step(np.zeros(8, np.uint8))
    assert o.shape == (8, 16 * 31) and o.sum() == 8 * 16 and ((r >= 0) & (r <= 1)).all()
