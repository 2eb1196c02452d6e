"""The reference's own tests/test_jax_game.py, re-stated against simulate_2048_b200.game
(torch CUDA tensors instead of jax arrays).  Test names and vectors follow the reference file
(line numbers cited) so that the two suites read side by side."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def T():
    import torch
    return torch


@pytest.fixture(scope='module')
def G():
    from simulate_2048_b200 import game
    return game


def arr(T, x, dtype=None):
    return T.tensor(x, dtype=dtype or T.int32, device='cuda')


class TestMergeRow:  # tests/test_jax_game.py:32-82
    @pytest.mark.parametrize('row,want,score', [
        ([0, 0, 0, 0], [0, 0, 0, 0], 0), ([0, 0, 2, 0], [2, 0, 0, 0], 0), ([2, 2, 0, 0], [4, 0, 0, 0], 4),
        ([2, 2, 2, 2], [4, 4, 0, 0], 8), ([2, 4, 8, 16], [2, 4, 8, 16], 0), ([2, 0, 2, 4], [4, 4, 0, 0], 4),
        ([4, 2, 2, 0], [4, 4, 0, 0], 4)])
    def test_rows(self, T, G, row, want, score):
        result, s = G.merge_row(arr(T, row))
        assert result.tolist() == want and int(s) == score


class TestSlideAndMerge:  # :85-95
    def test_full_board_slide(self, T, G):
        board = arr(T, [[2, 2, 4, 4], [4, 4, 2, 2], [2, 4, 2, 4], [4, 2, 4, 2]])
        result, score = G.slide_and_merge(board)
        assert result.tolist() == [[4, 8, 0, 0], [8, 4, 0, 0], [2, 4, 2, 4], [4, 2, 4, 2]]
        assert int(score) == 24


class TestLatentState:  # :98-135
    @pytest.mark.parametrize('board,action,want', [
        ([[0, 2, 0, 2], [0] * 4, [0] * 4, [0] * 4], 0, [[4, 0, 0, 0], [0] * 4, [0] * 4, [0] * 4]),
        ([[2, 0, 0, 0], [0] * 4, [2, 0, 0, 0], [0] * 4], 1, [[4, 0, 0, 0], [0] * 4, [0] * 4, [0] * 4]),
        ([[2, 0, 0, 2], [0] * 4, [0] * 4, [0] * 4], 2, [[0, 0, 0, 4], [0] * 4, [0] * 4, [0] * 4]),
        ([[2, 0, 0, 0], [0] * 4, [2, 0, 0, 0], [0] * 4], 3, [[0] * 4, [0] * 4, [0] * 4, [4, 0, 0, 0]])])
    def test_direction(self, T, G, board, action, want):
        result, reward = G.latent_state(arr(T, board), action)
        assert result.tolist() == want and int(reward) == 4


class TestLegalActionsMask:  # :138-163
    def test_empty_board_all_illegal(self, T, G):
        assert G.legal_actions_mask(T.zeros((4, 4), dtype=T.int32, device='cuda')).tolist() == [False] * 4

    def test_full_board_no_merges(self, T, G):
        board = arr(T, [[2, 4, 2, 4], [4, 2, 4, 2], [2, 4, 2, 4], [4, 2, 4, 2]])
        assert G.legal_actions_mask(board).tolist() == [False] * 4

    def test_can_slide_and_merge_left(self, T, G):
        assert bool(G.legal_actions_mask(arr(T, [[0, 2, 0, 0], [0] * 4, [0] * 4, [0] * 4]))[0])
        assert bool(G.legal_actions_mask(arr(T, [[2, 2, 0, 0], [0] * 4, [0] * 4, [0] * 4]))[0])


class TestIsDone:  # :166-182
    def test_cases(self, T, G):
        assert not bool(G.is_done(arr(T, [[2, 4, 0, 0], [0] * 4, [0] * 4, [0] * 4])))
        assert not bool(G.is_done(arr(T, [[2, 4, 2, 4], [4, 2, 4, 2], [2, 4, 2, 4], [4, 2, 2, 2]])))
        assert bool(G.is_done(arr(T, [[2, 4, 2, 4], [4, 2, 4, 2], [2, 4, 2, 4], [4, 2, 4, 2]])))


class TestFillCells:  # :185-206
    def test_fills_empty_cell(self, T, G):
        result = G.fill_cells(T.zeros((4, 4), dtype=T.int32, device='cuda'), 42)
        assert int((result != 0).sum()) == 1 and int(result.max()) in (2, 4)

    def test_full_board_unchanged(self, T, G):
        board = T.full((4, 4), 2, dtype=T.int32, device='cuda')
        assert T.equal(G.fill_cells(board, 42), board)


class TestNextState:  # :209-225
    def test_valid_move(self, T, G):
        board = arr(T, [[2, 0, 0, 0], [0] * 4, [2, 0, 0, 0], [0] * 4])
        result, reward = G.next_state(board, 1, 42)
        assert int(result[0, 0]) == 4 and int((result != 0).sum()) == 2 and float(reward) == 4.0
        assert reward.dtype == T.float32


class TestEnvironment:  # :228-270
    def test_reset(self, T, G):
        state = G.reset(42)
        assert isinstance(state, G.GameState) and state.tiles.shape == (4, 4)
        assert int((state.tiles != 0).sum()) == 2 and int(state.step_count) == 0
        assert not bool(state.done) and float(state.total_reward) == 0

    def test_step(self, T, G):
        state = G.reset(42)
        new_state, reward, done, info = G.step(state, 0, 43)
        assert int(new_state.step_count) == 1 and 'max_tile' in info

    def test_observation_and_mask(self, T, G):
        state = G.reset(42)
        obs, mask = G.get_observation(state), G.get_legal_actions(state)
        assert obs.shape == (16,) and obs.dtype == T.float32
        assert mask.shape == (4,) and mask.dtype == T.bool


class TestEncodeObservation:  # :273-291
    def test_known_values(self, T, G):
        assert float(G.encode_observation(T.zeros((4, 4), dtype=T.int32, device='cuda')).abs().sum()) == 0
        obs = G.encode_observation(arr(T, [[2, 4, 8, 16], [0] * 4, [0] * 4, [0] * 4]))
        assert obs[:4].tolist() == [1 / 16, 2 / 16, 3 / 16, 4 / 16]


class TestMaxTileCountEmpty:  # :294-317
    def test_values(self, T, G):
        board = arr(T, [[2, 4, 8, 16], [32, 64, 128, 256], [0] * 4, [0, 0, 0, 2048]])
        assert int(G.max_tile(board)) == 2048
        assert int(G.count_empty(T.zeros((4, 4), dtype=T.int32, device='cuda'))) == 16
        assert int(G.count_empty(arr(T, [[2, 0, 0, 0], [0, 4, 0, 0], [0] * 4, [0] * 4]))) == 14


class TestVectorization:  # :345-374 (vmap -> leading batch dimension)
    def test_batched_merge(self, T, G):
        rows = arr(T, [[2, 2, 0, 0], [4, 4, 0, 0], [2, 4, 2, 4]])
        results, scores = G.merge_row(rows)
        assert results.tolist() == [[4, 0, 0, 0], [8, 0, 0, 0], [2, 4, 2, 4]] and scores.tolist() == [4, 8, 0]

    def test_batched_reset(self, T, G):
        states = G.batched_reset(T.arange(10, dtype=T.int64, device='cuda') + 1000)
        assert states.tiles.shape == (10, 4, 4) and states.step_count.shape == (10,) and states.done.shape == (10,)
        assert ((states.tiles != 0).sum(dim=(1, 2)) == 2).all()


class TestBatchedEnvAgainstRestatement:
    """No reference test cross-checks the two engines; this one does: the batched
    API against the NumPy restatement of env.py:79-133 with injected spawns, incl. illegal moves,
    done pass-through, step_count and float32 total_reward."""

    def test_self_play_call_pattern(self, T, G, oracle):
        from oracle import np_engine as npe
        from tests.helpers import pack_tiles
        n, steps = 256, 120
        rng = np.random.default_rng(0)
        states = G.batched_reset(7, n)
        ref = [(t.astype(np.int64), 0, False, np.float32(0)) for t in states.tiles.cpu().numpy()]
        for _ in range(steps):
            obs = G.batched_get_observation(states)            # self_play.py:316
            legal = G.batched_get_legal_actions(states)        # self_play.py:317
            actions = rng.integers(0, 4, n)
            ranks, fours = np.zeros(n, np.int64), rng.random(n) < 0.1
            for i in range(n):
                after, _ = npe.jx_latent_state(ref[i][0], int(actions[i]))
                ranks[i] = rng.integers(0, 16) % max(int((after == 0).sum()), 1)
                np.testing.assert_array_equal(legal[i].cpu().numpy(), npe.jx_legal_actions_mask(ref[i][0]))
            spawns = T.tensor((ranks | np.where(fours, 0x10, 0)).astype(np.uint8), device='cuda')
            states, rewards, dones, info = G.batched_step(states, T.tensor(actions, device='cuda'), spawns)  # :354
            r_np, d_np = rewards.cpu().numpy(), dones.cpu().numpy()
            for i in range(n):
                st, r, d = npe.jx_env_step(*ref[i], int(actions[i]), int(ranks[i]), 4 if fours[i] else 2)
                ref[i] = st
                assert r == r_np[i] and d == d_np[i]
            np.testing.assert_array_equal(states.board.cpu().numpy().view(np.uint64),
                                          np.array([pack_tiles(s[0]) for s in ref], dtype=np.uint64))
            np.testing.assert_array_equal(states.step_count.cpu().numpy(), [s[1] for s in ref])
            np.testing.assert_array_equal(states.total_reward.cpu().numpy(), np.array([s[3] for s in ref], np.float32))
            np.testing.assert_array_equal(info['max_tile'].cpu().numpy(), [s[0].max() for s in ref])
            del obs
        assert states.done.any()


class TestSampleAction:
    """core.py:381 sample_action: the distribution is checked against the NumPy restatement
    (float32, tolerance 1e-6), the draw against that distribution."""

    CASES = [
        ([0.1, 0.2, 0.3, 0.4], [True, True, True, True]),
        ([0.1, 0.2, 0.3, 0.4], [True, False, True, False]),
        ([0.0, 0.0, 0.0, 1.0], [True, True, True, False]),   # all legal mass zero -> uniform over legal
        ([0.25, 0.25, 0.25, 0.25], [False, False, False, True]),
        ([0.7, 0.1, 0.1, 0.1], [False, False, False, False]),  # no legal action (finished game)
    ]

    @pytest.mark.parametrize('temperature', [1.0, 0.5, 0.25, 0.0])
    def test_distribution_matches_restatement(self, T, G, temperature):
        from oracle import np_engine as npe
        pol = T.tensor([c[0] for c in self.CASES], dtype=T.float32, device='cuda')
        legal = T.tensor([c[1] for c in self.CASES], device='cuda')
        got = G.core.action_distribution(temperature, pol, legal).cpu().numpy()
        for i, (p, m) in enumerate(self.CASES):
            want = npe.jx_action_distribution(temperature, np.array(p, np.float32), np.array(m))
            np.testing.assert_allclose(got[i], want, rtol=1e-5, atol=1e-6)

    def test_greedy_and_legal_only(self, T, G):
        pol = T.tensor([0.1, 0.6, 0.2, 0.1], device='cuda')
        legal = T.tensor([True, False, True, True], device='cuda')
        assert int(G.core.sample_action(0, 0.0, pol, legal)) == 2  # argmax among the legal actions
        draws = T.stack([G.core.sample_action(k, 1.0, pol, legal) for k in range(200)]).cpu().numpy()
        assert set(draws.tolist()) <= {0, 2, 3} and len(set(draws.tolist())) == 3

    def test_batched_draw_follows_distribution(self, T, G):
        n = 40000
        pol = T.tensor([0.1, 0.2, 0.3, 0.4], device='cuda').repeat(n, 1)
        legal = T.ones((n, 4), dtype=T.bool, device='cuda')
        picks = G.core.sample_action(7, 1.0, pol, legal).cpu().numpy()
        freq = np.bincount(picks, minlength=4) / n
        np.testing.assert_allclose(freq, [0.1, 0.2, 0.3, 0.4], atol=0.01)
