"""The reference's tests/test_board.py, test_move.py and test_environment.py re-stated against
simulate_2048_b200.twentyfortyeight (same names and vectors; ndarray in / ndarray out, rules
computed on the GPU)."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def gb():
    from simulate_2048_b200.twentyfortyeight.core import gameboard
    return gameboard


@pytest.fixture(scope='module')
def gm():
    from simulate_2048_b200.twentyfortyeight.core import gamemove
    return gamemove


@pytest.fixture(scope='module')
def Env():
    from simulate_2048_b200.twentyfortyeight.envs.twentyfortyeight import TwentyFortyEight
    return TwentyFortyEight


FULL = np.array([[2, 4, 8, 16], [32, 64, 128, 256], [512, 1024, 2048, 4096], [8192, 16384, 32768, 2]])


def test_reset(Env):  # test_board.py:33-37, test_environment.py:31-53
    env = Env(size=4)
    board = env.reset()
    assert np.count_nonzero(board) == 2 and np.all(np.isin(board[board != 0], [2, 4]))
    assert env.reward == 0 and not env.is_finished
    np.testing.assert_array_equal(env.reset(seed=42).copy(), env.reset(seed=42))


def test_step_valid_and_invalid_move(gb, gm, Env):  # test_board.py:39-51
    board = np.array([[2, 2, 0, 0], [0] * 4, [0] * 4, [0] * 4])
    nb, reward = gb.next_state(board, Env.ACTIONS['left'])
    assert nb[0, 0] == 4 and reward > 0
    board = np.array([[2, 0, 0, 0], [2, 0, 0, 0], [0] * 4, [0] * 4])
    nb, reward = gb.next_state(board, Env.ACTIONS['left'])
    assert nb is board and reward == 0  # gameboard.py:325: the same object comes back


def test_is_finished(gb):  # test_board.py:53-61 (the reference uses a 65536 tile; not representable)
    assert gb.is_done(FULL)
    assert not gb.is_done(np.array([[2, 0, 0, 0], [2, 0, 0, 0], [0] * 4, [0] * 4]))
    assert not gb.is_done(np.zeros((4, 4), dtype=np.int64))  # NumPy engine: empty board is not done


def test_merge_column_and_slide(gb):  # test_board.py:63-76, test_environment.py:187-219
    score, result = gb.merge_column(np.array([2, 2, 4, 4]))
    assert score == 12 and result.tolist() == [4, 8]
    score, result = gb.merge_column(np.array([0, 0, 0, 0]))
    assert score == 0 and len(result) == 0
    score, result = gb.merge_column(np.array([0, 4, 0, 0]))
    assert score == 0 and result.tolist() == [4]
    score, result = gb.merge_column(np.array([2, 2, 2, 2]))
    assert score == 8 and result.tolist() == [4, 4]
    score, result = gb.slide_and_merge(np.array([[2, 2, 4, 4], [0, 2, 2, 4], [2, 0, 0, 2], [2, 2, 2, 2]]))
    assert score == 28 and isinstance(score, float)
    assert result.tolist() == [[4, 8, 0, 0], [4, 4, 0, 0], [4, 0, 0, 0], [4, 4, 0, 0]]
    assert gb.slide_and_merge(np.array([[2, 2, 0, 0], [4, 4, 0, 0], [8, 8, 0, 0], [16, 16, 0, 0]]))[0] == 60


def test_after_state_lazy_equivalence(gb):  # test_board.py:78-113
    board = np.array([[2, 4, 0, 0], [0, 2, 0, 0], [0] * 4, [0, 0, 0, 2]])
    eager = gb.after_state(board)
    base, cells, num_empty = gb.after_state_lazy(board)
    lazy = [gb.generate_outcome(base, c, v, num_empty) for c in cells for v in (2, 4)]
    assert len(eager) == len(lazy) == 2 * num_empty
    for (e, pe), (l, pl) in zip(eager, lazy):  # same order, too
        np.testing.assert_array_equal(e, l)
        assert abs(pe - pl) < 1e-10
    base, cells, num_empty = gb.after_state_lazy(FULL)
    assert num_empty == 0 and cells == []
    with pytest.raises(ValueError):
        gb.generate_outcome(FULL, (0, 0), 2, 0)


def test_legal_actions(gm):  # test_move.py:9-23, test_environment.py:166-181
    board = np.array([[2, 0, 0, 0], [2, 0, 0, 0], [0] * 4, [0] * 4])
    assert set(gm.illegal_actions(board)) == {0} and set(gm.legal_actions(board)) == {1, 2, 3}
    assert gm.legal_actions_mask(board) == (False, True, True, True)
    b = np.zeros((4, 4), dtype=int); b[0, 0] = 2; b[1, 1] = 2
    assert len(gm.legal_actions(b)) == 4
    edge = np.array([[2, 0, 0, 0], [4, 0, 0, 0], [8, 0, 0, 0], [16, 0, 0, 0]])
    assert 0 in gm.illegal_actions(edge) and not gm.can_move(edge) and gm._can_move_direction(edge, 2)


def test_invalid_move_no_state_change(gb):  # test_environment.py:121-149
    board = np.array([[2, 0, 0, 0], [4, 0, 0, 0], [8, 0, 0, 0], [16, 0, 0, 0]])
    nb, reward = gb.next_state(board, 0)
    np.testing.assert_array_equal(nb, board)
    assert reward == 0
    nb, reward = gb.next_state(np.array([[2, 2, 0, 0], [0] * 4, [0] * 4, [0] * 4]), 0, seed=42)
    assert reward == 4 and np.count_nonzero(nb) == 2


def test_termination_boards(gb):  # test_environment.py:225-243
    assert gb.is_done(FULL)
    assert not gb.is_done(np.array([[2, 4, 8, 0], [16, 32, 64, 128], [256, 512, 1024, 2048], [4096, 8192, 16384, 32768]]))
    assert not gb.is_done(np.array([[2, 2, 4, 8], [16, 32, 64, 128], [256, 512, 1024, 2048], [4096, 8192, 16384, 32768]]))


def test_spawn_distribution(gb, Env):  # test_environment.py:249-302
    counts = {2: 0, 4: 0}
    for i in range(300):
        env = Env(size=4)
        env.reset(seed=i)
        for t in env._current_state[env._current_state != 0]:
            counts[int(t)] += 1
    assert abs(counts[4] / sum(counts.values()) - 0.1) < 0.05
    outcomes = gb.after_state(np.array([[2, 4, 0, 0], [0] * 4, [0] * 4, [0] * 4]))
    assert abs(sum(p for _, p in outcomes) - 1.0) < 1e-10
    one = np.array([[2, 4, 8, 16], [32, 64, 128, 0], [256, 512, 1024, 2048], [4096, 8192, 16384, 32768]])
    probs = sorted(p for _, p in gb.after_state(one))
    assert len(probs) == 2 and abs(probs[0] - 0.1) < 1e-10 and abs(probs[1] - 0.9) < 1e-10


def test_fill_cells_in_place(gb):
    z = np.zeros((4, 4), dtype=np.int64)
    out = gb.fill_cells(z, 2, seed=3)
    assert out is z and np.count_nonzero(z) == 2
    full = np.full((4, 4), 2, dtype=np.int64)
    assert gb.fill_cells(full, 1).tolist() == [[2] * 4] * 4


def test_game_reaches_termination(gm, Env):  # test_environment.py:308-328
    env = Env(size=4)
    env.reset(seed=42)
    steps = 0
    while steps < 2000:
        legal = gm.legal_actions(env._current_state)
        if not legal or env.is_finished:
            break
        _, _, done = env.step(legal[0])
        steps += 1
        if done:
            break
    assert env.is_finished and steps < 2000


def test_out_of_scope_variants_raise(Env):
    with pytest.raises(NotImplementedError):
        Env(size=6)


def test_against_oracle_random_boards(gb, gm, oracle):
    from tests.helpers import synthetic_packed
    boards = synthetic_packed(200, seed=99, max_exp=5)
    tiles = oracle.unpack(boards).astype(np.int64)
    for t in tiles[:60]:
        assert sum(int(v) << a for a, v in enumerate(gm.legal_actions_mask(t))) == oracle.legal_mask_values(t)
        for a in range(4):
            s, r = gb.latent_state(t, a)
            os_, or_ = oracle.latent_state_values(t, a)
            np.testing.assert_array_equal(s, os_)
            assert r == or_


def test_encoded_and_normalized_variants(Env, golden):
    """TwentyFortyEight(encoded=True / normalize=True) (twentyfortyeight.py:66-67, :80-81) against
    fixtures from the reference's utils/binary.py and utils/normalize.py."""
    import torch
    from simulate_2048_b200 import ops
    from simulate_2048_b200.twentyfortyeight.utils.binary import encode_flatten
    from simulate_2048_b200.twentyfortyeight.utils.normalize import normalize_reward
    from tests.helpers import pack_tiles
    g = golden('encodings')
    onehot = ops.observe_onehot(torch.from_numpy(pack_tiles(g['tiles']).view(np.int64)).cuda(), 31)
    np.testing.assert_array_equal(onehot.reshape(len(g['tiles']), -1).cpu().numpy(), g['onehot'])
    enc = encode_flatten(g['tiles'][3], encodage_size=31)
    assert enc.dtype == np.int64 and enc.shape == (496,)
    np.testing.assert_array_equal(enc, g['onehot'][3])
    for r, want in zip(g['rewards'], g['normalized']):
        assert normalize_reward(float(r)) == want
    # test_environment.py:87-97: encoded observation length 4*4*31
    env = Env(size=4, encoded=True)
    assert env.reset().ndim == 1 and len(env.observation) == 4 * 4 * 31
    # test_environment.py:99-115: normalised reward is log-compressed
    env_n, env_r = Env(size=4, normalize=True), Env(size=4)
    board = np.array([[2, 2, 0, 0], [0] * 4, [0] * 4, [0] * 4])
    env_n._current_state, env_r._current_state = board.copy(), board.copy()
    _, rn, _ = env_n.step(0)
    _, rr, _ = env_r.step(0)
    assert 0 < rn < rr and rn == 2 / 16
