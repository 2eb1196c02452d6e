"""Pins the CPU oracle (oracle/g2048_oracle.c, oracle/np_engine.py) BEFORE it is trusted:

  * against tests/golden/*.npz, produced by the real reference NumPy engine (make_golden.py);
  * against the known-answer vectors in the reference's own tests (cited file:line);
  * against the SHA-256 hashes recorded in SURVEY.md §8c;
  * C oracle and NumPy restatement against each other, and — when /root/reference is present
    (the build container) — against the imported reference itself.
"""

import hashlib
import os
import sys

import numpy as np
import pytest

from oracle import np_engine as npe
from tests.helpers import pack_tiles, synthetic_packed

FLAG_DONE, FLAG_MOVED, FLAG_OVERFLOW, FLAG_BADSPAWN = 0x10, 0x20, 0x40, 0x80


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


# ---------------------------------------------------------------- row table

def test_row_table_matches_reference_merge_column(oracle, golden):
    g = golden('rows')
    res, rew, mv = oracle.row_table()
    np.testing.assert_array_equal(res, g['result'])
    np.testing.assert_array_equal(rew, g['score'])
    assert sha(res) == '2311124775fc20afc51de212e97fc3f04eef16691d5d270594080c77806f1cac'
    assert sha(rew) == '5998c5d66dcf2d4a7725a856533105ebffeff7fd3ef1825833cbbe9fc6bc9949'
    assert sha(mv) == '81e7080ebdd8dbbd6331c30b323f7746193fc5f48e11eb5428402cf7c4f3e30d'
    assert int(mv.sum()) == 21210 and int(rew.sum()) == 100660224 and int(rew.max()) == 131072
    for row, out, score in [(0x0011, 0x0002, 4), (0x1111, 0x0022, 8), (0x2110, 0x0022, 4),
                            (0x1022, 0x0013, 8), (0x0121, 0x0121, 0), (0x4321, 0x4321, 0)]:
        assert (res[row], rew[row]) == (out, score)


def test_row_overflow_rows(oracle):
    res, rew, _ = oracle.row_table()
    ovf = 0
    for r in range(65536):
        cells = [(r >> (4 * j)) & 0xF for j in range(4) if (r >> (4 * j)) & 0xF]
        i, hit = 0, False
        while i < len(cells) - 1:
            if cells[i] == cells[i + 1]:
                hit |= cells[i] == 15
                i += 2
            else:
                i += 1
        ovf += hit
    assert ovf == 767  # SURVEY.md §7 "Hard parts"


# ---------------------------------------------------------------- Philox known answers

def test_philox_known_answers(oracle):
    kat = [
        ([0, 0, 0, 0], [0, 0], [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]),
        ([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2, [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]),
        ([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0],
         [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]),
    ]
    for ctr, key, out in kat:
        assert oracle.philox4x32_10(ctr, key).tolist() == out


def _philox_py(ctr, key):
    """Independent big-int Philox4x32-10 (Salmon et al. 2011), checked by the KATs above."""
    c, k = list(ctr), list(key)
    for _ in range(10):
        p0, p1 = 0xD2511F53 * c[0], 0xCD9E8D57 * c[2]
        c = [(p1 >> 32) ^ c[1] ^ k[0], p1 & 0xFFFFFFFF, (p0 >> 32) ^ c[3] ^ k[1], p0 & 0xFFFFFFFF]
        k = [(k[0] + 0x9E3779B9) & 0xFFFFFFFF, (k[1] + 0xBB67AE85) & 0xFFFFFFFF]
    return c


def test_philox_stream_layout_v2(oracle):
    """include/g2048.h "Philox mode": one block per env quad, env e reads word e & 3; the word is a
    mixed-radix fraction (rank base n_empty, tile-4 digit base 10, remainder)."""
    rng = np.random.default_rng(5)
    for _ in range(200):
        key = int(rng.integers(0, 2**63)) * 2 + int(rng.integers(0, 2))
        env = int(rng.integers(0, 2**40))
        step, tag = int(rng.integers(0, 2**32)), int(rng.integers(0, 4))
        q = env >> 2
        block = _philox_py([q & 0xFFFFFFFF, q >> 32, step, tag], [key & 0xFFFFFFFF, key >> 32])
        assert _philox_py([0, 0, 0, 0], [0, 0]) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
        w = oracle.philox_word(key, env, step, tag)
        assert w == block[env & 3]
        n = int(rng.integers(1, 17))
        rank, four, rem = oracle.draw_spawn(w, n)
        assert rank == (w * n) >> 32 and 0 <= rank < n
        f1 = (w * n) & 0xFFFFFFFF
        assert four == ((f1 * 10) >> 32 == 0) and rem == (f1 * 10) & 0xFFFFFFFF
    # pinned words of the v2 stream (derived with the KAT-checked generator above)
    assert [oracle.philox_word(0, e, 0, 0) for e in range(4)] == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert oracle.philox_word(2048, 5, 7, 0) == _philox_py([1, 0, 7, 0], [2048, 0])[1]


def test_game_stream_layout_v3(oracle):
    """include/g2048.h "Game stream" (g2048_reset, g2048_rollout): word p of env e = word p & 3 of
    the block of counter (e_lo, e_hi, p >> 2, tag 6) — checked against the independent generator; the
    reset tiles are words 0 and 1, and a from-reset rollout reads word step + 2."""
    rng = np.random.default_rng(6)
    for _ in range(200):
        key = int(rng.integers(0, 2**63)) * 2 + int(rng.integers(0, 2))
        env = int(rng.integers(0, 2**40))
        p = int(rng.integers(0, 2**20))
        block = _philox_py([env & 0xFFFFFFFF, env >> 32, p >> 2, 6], [key & 0xFFFFFFFF, key >> 32])
        assert oracle.game_word(key, env, p) == block[p & 3]
    assert [oracle.game_word(0, 0, p) for p in range(4)] == _philox_py([0, 0, 0, 6], [0, 0])
    # reset = two spawns from words 0, 1 (rank base 16, then base 15; tile 4 iff digit 2 == 0)
    for env in (0, 1, 5, 10**12 + 3):
        tiles = [0] * 16
        for t in range(2):
            empties = [c for c in range(16) if tiles[c] == 0]
            rank, four, _ = oracle.draw_spawn(oracle.game_word(77, env, t), len(empties))
            tiles[empties[rank]] = 4 if four else 2
        got = oracle.unpack(oracle.reset(1, seed=77, env_offset=env)).reshape(-1).tolist()
        assert got == tiles
    # a from-reset rollout's first spawn is word 2: replay one step by hand
    out = oracle.rollout(4, seed=9, trace_stride=64)
    for i in range(4):
        byte = int(out['trace'][i, 0])
        b0 = oracle.reset(1, seed=9, env_offset=i)
        after = oracle.latent_state(int(b0[0]), byte & 3)[0]
        empties = [c for c in range(16) if (after >> (4 * c)) & 0xF == 0]
        rank, four, _ = oracle.draw_spawn(oracle.game_word(9, i, 2), len(empties))
        assert empties[rank] == (byte >> 2) & 0xF and bool(four) == bool(byte & 0x40)


def test_rollout_oracle_against_an_independent_replay(oracle):
    """orc_rollout (the C oracle the GPU rollouts are compared with) replayed WITHOUT any C code: the
    NumPy restatement of the reference engine for the transitions, the big-int Philox above for the
    game stream, the mixed-radix digits and the k-th-legal-action policy in plain Python — whole
    games from reset (words 0, 1 = reset tiles, word step + 2 per step) and from given boards (first
    action from the quad-layout word of tag 3, word step per step, a step cap)."""
    from oracle import np_engine as ne

    def word(key, env, p):
        return _philox_py([env & 0xFFFFFFFF, env >> 32, p >> 2, 6], [key & 0xFFFFFFFF, key >> 32])[p & 3]

    def spawn(board, w):
        free = np.flatnonzero(board.reshape(-1) == 0)
        p1 = w * len(free)
        p2 = (p1 & 0xFFFFFFFF) * 10
        out = board.reshape(-1).copy()
        out[free[p1 >> 32]] = 4 if (p2 >> 32) == 0 else 2
        return out.reshape(4, 4), p2 & 0xFFFFFFFF

    def play(board, rem, key, env, first_word, cap):
        score = steps = 0
        while steps < cap:
            legal = ne.legal_actions(board)
            if not legal:
                break
            action = legal[(rem * len(legal)) >> 32]
            after, reward = ne.latent_state(board, action)
            board, rem = spawn(np.asarray(after), word(key, env, first_word + steps))
            score += int(reward)
            steps += 1
        return board, score, steps

    key, off = 0x1234567890ABCDEF, 10**12 + 5
    want = oracle.rollout(6, seed=key, env_offset=off)
    for i in range(6):
        env = off + i
        b = np.zeros((4, 4), dtype=np.int64)
        b, _ = spawn(b, word(key, env, 0))
        b, rem = spawn(b, word(key, env, 1))
        b, score, steps = play(b, rem, key, env, 2, 1 << 20)
        assert steps == want['steps'][i] and score == want['score'][i]
        assert int(oracle.pack(b.reshape(1, 16).astype(np.int32))[0][0]) == int(want['boards'][i])
    start = oracle.reset(6, seed=3)
    want = oracle.rollout(6, seed=key, env_offset=off, from_reset=False, boards=start, max_steps=40)
    for i in range(6):
        env = off + i
        q = env >> 2
        rem = _philox_py([q & 0xFFFFFFFF, q >> 32, 0, 3], [key & 0xFFFFFFFF, key >> 32])[env & 3]
        b = oracle.unpack(start[i:i + 1]).reshape(4, 4).astype(np.int64)
        b, score, steps = play(b, rem, key, env, 0, 40)
        assert steps == want['steps'][i] and score == want['score'][i]
        assert int(oracle.pack(b.reshape(1, 16).astype(np.int32))[0][0]) == int(want['boards'][i])


def test_philox_digits_are_uniform(oracle):
    """Digit 2 (p = 0.1) and digit 3 are independent of digit 1: chi-square-free sanity bounds."""
    n = 200_000
    ws = [oracle.philox_word(99, e, 3, 0) for e in range(0, n)]
    ranks = np.zeros(7, dtype=np.int64)
    fours = np.zeros(7, dtype=np.int64)
    acts = np.zeros((2, 3), dtype=np.int64)
    for w in ws:
        r, four, rem = oracle.draw_spawn(w, 7)
        ranks[r] += 1
        fours[r] += four
        acts[int(four), (rem * 3) >> 32] += 1
    assert np.all(np.abs(ranks / n - 1 / 7) < 0.004)
    assert np.all(np.abs(fours / ranks - 0.1) < 0.008)          # P(4 | rank) = 0.1 for every rank
    assert np.all(np.abs(acts / acts.sum(axis=1, keepdims=True) - 1 / 3) < 0.012)  # action digit given the tile


# ---------------------------------------------------------------- reference test vectors

def test_reference_known_answers_numpy_engine(oracle):
    # tests/test_board.py:63-68
    out, score = oracle.merge_line_values([2, 2, 4, 4])
    assert score == 12 and out.tolist() == [4, 8, 0, 0]
    # tests/test_environment.py:187-210
    assert oracle.merge_line_values([0, 0, 0, 0])[1] == 0
    assert oracle.merge_line_values([0, 4, 0, 0])[0].tolist() == [4, 0, 0, 0]
    out, score = oracle.merge_line_values([2, 2, 2, 2])
    assert score == 8 and out.tolist() == [4, 4, 0, 0]
    # tests/test_board.py:70-76
    board = [[2, 2, 4, 4], [0, 2, 2, 4], [2, 0, 0, 2], [2, 2, 2, 2]]
    after, score = oracle.latent_state_values(board, 0)
    assert score == 28
    assert after.tolist() == [[4, 8, 0, 0], [4, 4, 0, 0], [4, 0, 0, 0], [4, 4, 0, 0]]
    # tests/test_environment.py:212-219
    assert oracle.latent_state_values([[2, 2, 0, 0], [4, 4, 0, 0], [8, 8, 0, 0], [16, 16, 0, 0]], 0)[1] == 60
    # tests/test_move.py:9-23
    assert oracle.legal_mask_values([[2, 0, 0, 0], [2, 0, 0, 0], [0] * 4, [0] * 4]) == 0b1110
    # tests/test_environment.py:166-181
    b = np.zeros((4, 4), dtype=np.int64); b[0, 0] = 2; b[1, 1] = 2
    assert oracle.legal_mask_values(b) == 0b1111
    assert oracle.legal_mask_values([[2, 0, 0, 0], [4, 0, 0, 0], [8, 0, 0, 0], [16, 0, 0, 0]]) & 1 == 0
    # tests/test_board.py:53-61, tests/test_environment.py:225-243 (a 65536 tile: values only)
    full = [[2, 4, 8, 16], [32, 64, 128, 256], [512, 1024, 2048, 4096], [8192, 16384, 32768, 65536]]
    assert oracle.is_done_numpy_values(full)
    assert not oracle.is_done_numpy_values([[2, 0, 0, 0], [2, 0, 0, 0], [0] * 4, [0] * 4])
    assert not oracle.is_done_numpy_values([[2, 4, 8, 0], [16, 32, 64, 128], [256, 512, 1024, 2048], [4096, 8192, 16384, 32768]])
    assert not oracle.is_done_numpy_values([[2, 2, 4, 8], [16, 32, 64, 128], [256, 512, 1024, 2048], [4096, 8192, 16384, 32768]])


def test_reference_known_answers_jax_engine(oracle):
    """Literal vectors of tests/test_jax_game.py — C oracle and the NumPy restatement of the JAX
    semantics must both reproduce them."""
    rows = [  # tests/test_jax_game.py:35-82
        ([0, 0, 0, 0], [0, 0, 0, 0], 0), ([0, 0, 2, 0], [2, 0, 0, 0], 0), ([2, 2, 0, 0], [4, 0, 0, 0], 4),
        ([2, 2, 2, 2], [4, 4, 0, 0], 8), ([2, 4, 8, 16], [2, 4, 8, 16], 0), ([2, 0, 2, 4], [4, 4, 0, 0], 4),
        ([4, 2, 2, 0], [4, 4, 0, 0], 4),
    ]
    for row, want, score in rows:
        out, s = oracle.merge_line_values(row)
        assert (out.tolist(), s) == (want, score)
        out, s = npe.jx_merge_row(np.array(row))
        assert (out.tolist(), s) == (want, score)
    # :88-95
    board = np.array([[2, 2, 4, 4], [4, 4, 2, 2], [2, 4, 2, 4], [4, 2, 4, 2]])
    want = [[4, 8, 0, 0], [8, 4, 0, 0], [2, 4, 2, 4], [4, 2, 4, 2]]
    for fn in (oracle.latent_state_values, npe.jx_latent_state):
        after, score = fn(board, 0)
        assert np.asarray(after).tolist() == want and score == 24
    # :101-135 one golden per direction
    cases = [
        ([[0, 2, 0, 2], [0] * 4, [0] * 4, [0] * 4], 0, [[4, 0, 0, 0], [0] * 4, [0] * 4, [0] * 4]),
        ([[2, 0, 0, 0], [0] * 4, [2, 0, 0, 0], [0] * 4], 1, [[4, 0, 0, 0], [0] * 4, [0] * 4, [0] * 4]),
        ([[2, 0, 0, 2], [0] * 4, [0] * 4, [0] * 4], 2, [[0, 0, 0, 4], [0] * 4, [0] * 4, [0] * 4]),
        ([[2, 0, 0, 0], [0] * 4, [2, 0, 0, 0], [0] * 4], 3, [[0] * 4, [0] * 4, [0] * 4, [4, 0, 0, 0]]),
    ]
    for b, a, want in cases:
        for fn in (oracle.latent_state_values, npe.jx_latent_state):
            after, score = fn(np.array(b), a)
            assert np.asarray(after).tolist() == want and score == 4
    # :141-151 empty board and checkerboard: nothing legal; :169-182 is_done
    checker = np.array([[2, 4, 2, 4], [4, 2, 4, 2], [2, 4, 2, 4], [4, 2, 4, 2]])
    assert oracle.legal_mask_values(np.zeros((4, 4))) == 0 and not npe.jx_legal_actions_mask(np.zeros((4, 4), dtype=int)).any()
    assert oracle.legal_mask_values(checker) == 0 and npe.jx_is_done(checker)
    assert oracle.mask(pack_tiles(checker).reshape(1))[0] == FLAG_DONE
    assert not npe.jx_is_done(np.array([[2, 4, 0, 0], [0] * 4, [0] * 4, [0] * 4]))
    assert not npe.jx_is_done(np.array([[2, 4, 2, 4], [4, 2, 4, 2], [2, 4, 2, 4], [4, 2, 2, 2]]))
    # :282-291 observation
    obs, mx = oracle.observe(pack_tiles(np.array([[2, 4, 8, 16], [0] * 4, [0] * 4, [0] * 4])).reshape(1))
    assert obs[0, :4].tolist() == [1 / 16, 2 / 16, 3 / 16, 4 / 16] and mx[0] == 4
    np.testing.assert_array_equal(obs[0], npe.jx_encode_observation(np.array([[2, 4, 8, 16], [0] * 4, [0] * 4, [0] * 4])))


# ---------------------------------------------------------------- golden: synthetic set

def test_synthetic_set_hashes(golden):
    g = golden('synthetic')
    assert sha(g['packed']) == '0a59753ef220a7566a61001718f751ec1353da84d5c94cf4c6c7510e0dd5efde'
    assert sha(g['after']) == '0fa1c74c638f8fa50c18ee39b115249a5a9b5e30f37fe2dd2364c4105802538b'
    assert sha(g['rewards']) == '9186946ed0fbe81d4d5a329a37cb0cca0265d244c6165be631b5fe2effe7f7b8'
    assert sha(g['mask']) == '4174e6e66051f557078e3e4033bce1a4c83f725e5ae5cb71b1c2b94519f478c8'
    assert sha(g['done']) == '9322adcd96ae75c9949623b0a84aa52c26ba9dd7a88e61100ce45633ba02985a'
    assert sha(g['child_count']) == 'ed5cbedb117fd1121a2477af36ae2741b5f6711f2b28642cf29e1bdf84c1a9e8'
    np.testing.assert_array_equal(g['packed'], synthetic_packed(4096))


def test_oracle_expand_matches_reference(oracle, golden):
    g = golden('synthetic')
    ex = oracle.expand(g['packed'], child_layout=2)
    np.testing.assert_array_equal(ex['after'], g['after'])
    np.testing.assert_array_equal(ex['rewards'], g['rewards'].astype(np.int32))
    np.testing.assert_array_equal(ex['mask'] & 0xF, g['mask'])
    np.testing.assert_array_equal((ex['mask'] & FLAG_DONE) != 0, g['mask'] == 0)
    np.testing.assert_array_equal(ex['child_count'], g['child_count'])
    np.testing.assert_array_equal(ex['children'], g['children'])  # order: action, cell, value
    # probabilities are implied: {0.9, 0.1} / num_empty (gameboard.py:166)
    ne = np.repeat(ex['num_empty'].reshape(-1), ex['child_count'].reshape(-1)).astype(np.float64)
    probs = np.where(np.arange(ne.size) % 2 == 0, 0.9, 0.1) / ne
    np.testing.assert_allclose(probs, g['child_probs'], rtol=0, atol=1e-15)
    # dense layout carries the same children
    dense = oracle.expand(g['packed'], child_layout=1)['children'].reshape(-1, 32)
    cc = g['child_count'].reshape(-1)
    flat = np.concatenate([dense[i, : cc[i]] for i in range(cc.size)])
    np.testing.assert_array_equal(flat, g['children'])


def test_numpy_is_done_vs_mask(golden):
    g = golden('synthetic')
    # gameboard.py:328 is_done == "no legal action" on every non-empty board
    np.testing.assert_array_equal(g['done'] != 0, g['mask'] == 0)


def test_oracle_step_matches_reference_next_state(oracle, golden):
    g = golden('synthetic')
    legal = g['next_spawn'] != 0xFF
    spawns = np.where(legal, g['next_spawn'], 0).astype(np.uint8)
    out, rew, fl = oracle.step(g['packed'], g['actions'], spawns)
    np.testing.assert_array_equal(out, g['next'])
    np.testing.assert_array_equal(rew, g['next_reward'].astype(np.int32))
    np.testing.assert_array_equal((fl & FLAG_MOVED) != 0, legal)
    assert not (fl & (FLAG_OVERFLOW | FLAG_BADSPAWN)).any()
    # flags describe the NEW board
    np.testing.assert_array_equal(fl & 0x1F, oracle.mask(out))


def test_oracle_replays_reference_games(oracle, golden):
    g = golden('games')
    legal = g['spawn'] != 0xFF
    spawns = np.where(legal, g['spawn'], 0).astype(np.uint8)
    out, rew, fl = oracle.step(g['board'], g['action'], spawns)
    np.testing.assert_array_equal(out, g['next'])
    np.testing.assert_array_equal(rew, g['reward'].astype(np.int32))
    np.testing.assert_array_equal((fl & FLAG_DONE) != 0, g['done'] != 0)
    np.testing.assert_array_equal((fl & FLAG_MOVED) != 0, legal)
    # trajectories chain: the next board of step t is the board of step t+1 within a game
    same = g['game'][1:] == g['game'][:-1]
    np.testing.assert_array_equal(g['next'][:-1][same], g['board'][1:][same])
    # reset boards: exactly two tiles in {2, 4}
    for b in g['first']:
        e = [(int(b) >> (4 * i)) & 0xF for i in range(16)]
        assert sorted(x for x in e if x)[0] in (1, 2) and sum(1 for x in e if x) == 2


def test_oracle_spawn_matches_reference_fill_cells(oracle, golden):
    g = golden('fill_cells')
    base = np.uint64(g['base'])
    for cell, value, rank in g['draws'].tolist():
        byte = np.array([rank | (0x10 if value == 4 else 0)], dtype=np.uint8)
        got, fl = oracle.spawn(np.array([base]), byte)
        want = int(base) | ((2 if value == 4 else 1) << (4 * cell))
        assert int(got[0]) == want and fl[0] == 0
        got2, _ = oracle.spawn(np.array([base]), np.array([0x80 | cell | (0x10 if value == 4 else 0)], dtype=np.uint8))
        assert int(got2[0]) == want
    assert g['draws'][:8].tolist() == [[8, 2, 6], [11, 2, 9], [2, 2, 1], [3, 2, 2], [14, 4, 11], [1, 2, 0], [8, 2, 6], [10, 2, 8]]


# ---------------------------------------------------------------- oracle self-consistency

def test_c_oracle_vs_numpy_restatement(oracle):
    rng = np.random.default_rng(11)
    boards = synthetic_packed(600, seed=5, max_exp=4)  # small exponents force merges, full boards
    tiles = oracle.unpack(boards).astype(np.int64)
    for i in range(len(boards)):
        t = tiles[i]
        mask = npe.legal_actions_mask(t)
        assert oracle.legal_mask_values(t) == sum(int(v) << a for a, v in enumerate(mask))
        np.testing.assert_array_equal(npe.jx_legal_actions_mask(t), np.array(mask))
        assert npe.is_done(t) == oracle.is_done_numpy_values(t)
        for a in range(4):
            a1, r1 = npe.latent_state(t, a)
            a2, r2 = npe.jx_latent_state(t, a)
            a3, r3 = oracle.latent_state_values(t, a)
            np.testing.assert_array_equal(a1, a3)
            np.testing.assert_array_equal(a2, a3)
            assert r1 == r3 == r2
    del rng


def test_oracle_env_step_follows_jax_semantics(oracle):
    """env.py:79-133 through the C oracle vs the NumPy restatement, incl. done pass-through and
    step_count incrementing on illegal moves."""
    n, steps = 64, 60
    rng = np.random.default_rng(3)
    boards = oracle.reset(n, seed=9)
    state = (boards, np.zeros(n, np.int32), np.zeros(n, np.uint8), np.zeros(n, np.float32))
    ref = [(oracle.unpack(boards[i:i + 1])[0].astype(np.int64), 0, False, np.float32(0)) for i in range(n)]
    for _ in range(steps):
        actions = rng.integers(0, 4, n).astype(np.uint8)
        ranks = rng.integers(0, 16, n)
        fours = rng.random(n) < 0.1
        # valid ranks need the afterstate's empties
        spawns = np.zeros(n, np.uint8)
        for i in range(n):
            after, _ = npe.jx_latent_state(ref[i][0], int(actions[i]))
            ne = int((after == 0).sum())
            ranks[i] = ranks[i] % max(ne, 1)
            spawns[i] = ranks[i] | (0x10 if fours[i] else 0)
        state, rew, fl, obs, mt = oracle.env_step(*state, actions, spawns)
        for i in range(n):
            st, r, d = npe.jx_env_step(*ref[i], int(actions[i]), int(ranks[i]), 4 if fours[i] else 2)
            ref[i] = st
            assert pack_tiles(st[0]) == state[0][i]
            assert st[1] == state[1][i] and bool(st[2]) == bool(state[2][i])
            assert st[3] == state[3][i] and r == rew[i] and d == bool(fl[i] & FLAG_DONE or state[2][i])
            np.testing.assert_array_equal(obs[i], npe.jx_encode_observation(st[0]))
            assert mt[i] == st[0].max()


def test_oracle_overflow_and_bad_spawn(oracle):
    b = pack_tiles(np.array([[32768, 32768, 0, 0], [0] * 4, [0] * 4, [0] * 4])).reshape(1)
    out, rew, fl = oracle.step(b, np.array([0], np.uint8), np.array([0], np.uint8))
    assert fl[0] & FLAG_OVERFLOW and rew[0] == 65536
    assert int(out[0]) & 0xF == 1  # wrapped merged nibble 0, first empty cell got the 2
    b = pack_tiles(np.array([[2, 2, 0, 0], [0] * 4, [0] * 4, [0] * 4])).reshape(1)
    _, _, fl = oracle.step(b, np.array([0], np.uint8), np.array([15], np.uint8))  # rank 15 of 15 empties
    assert fl[0] & FLAG_BADSPAWN
    _, _, fl = oracle.step(b, np.array([0], np.uint8), np.array([0x80], np.uint8))  # cell 0 occupied
    assert fl[0] & FLAG_BADSPAWN


def test_oracle_rollout_trace_replays(oracle):
    out = oracle.rollout(32, seed=5, trace_stride=2048)
    for i in range(32):
        b = oracle.reset(1, seed=5, env_offset=i)
        score = 0
        t = 0
        while out['trace'][i, t]:
            byte = int(out['trace'][i, t])
            spawn = np.array([0x80 | ((byte >> 2) & 0xF) | (0x10 if byte & 0x40 else 0)], np.uint8)
            b, r, fl = oracle.step(b, np.array([byte & 3], np.uint8), spawn)
            assert fl[0] & FLAG_MOVED
            score += int(r[0])
            t += 1
        assert t == out['steps'][i] and score == out['score'][i] and b[0] == out['boards'][i]
        assert fl[0] & FLAG_DONE
    st = out['stats']
    assert st[0] == 32 and st[1] == out['steps'].sum() and st[2] == out['score'].sum()
    assert st[3] == out['score'].max() and st[4] == out['steps'].max() and st[8:].sum() == 32


def test_oracle_philox_spawn_distribution(oracle):
    """tests/test_environment.py:249-279 — tiles in {2,4}, 90/10 within ±5 % — on the Philox reset."""
    boards = oracle.reset(20000, seed=123)
    tiles = oracle.unpack(boards).reshape(-1)
    nz = tiles[tiles != 0]
    assert nz.size == 40000 and set(np.unique(nz)) == {2, 4}
    assert abs((nz == 4).mean() - 0.1) < 0.01
    cells = np.bincount(np.nonzero(oracle.unpack(boards).reshape(-1, 16))[1], minlength=16) / 40000
    assert np.abs(cells - 1 / 16).max() < 0.01


def test_rollout_stats_fixture_is_the_oracles(oracle):
    """tests/golden/rollout_stats.json (what bench.py checks the sharded 2^24-env GPU run against) is
    the oracle's own output: recomputed here for the 65536-env entry, both the threaded driver
    that generated the file and the plain single-thread loop."""
    import ctypes as C
    import json
    fix = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'rollout_stats.json')))
    assert set(fix['stats']) >= {'65536', '1048576', '2097152', '16777216'} and fix['seed'] == 2048
    for n, words in fix['stats'].items():
        assert words[0] == int(n) and sum(words[8:]) == int(n) and 110 < words[1] / words[0] < 127
    lib = oracle.lib()
    stats = np.zeros(24, dtype=np.uint64)
    lib.orc_set_num_threads(4)
    lib.orc_rollout_mt(C.c_uint64(2048), C.c_uint64(0), C.c_size_t(65536), C.c_uint32(1 << 20),
                       stats.ctypes.data_as(C.c_void_p))
    assert stats.tolist() == fix['stats']['65536']
    part = oracle.rollout(4096, seed=2048)
    assert int(part['stats'][0]) == 4096


# ---------------------------------------------------------------- live reference (build box only)

def _staged_reference():
    from oracle import make_ref
    return make_ref.available()


@pytest.mark.skipif(not _staged_reference(), reason='oracle/_ref not staged (python -m oracle.make_ref)')
def test_restatement_against_live_reference(oracle):
    """The restatements against the REAL reference engine, imported from oracle/_ref (a verbatim
    copy of /root/reference/twentyfortyeight made by oracle/make_ref.py; it ships to the GPU box)."""
    from oracle import make_ref
    ref, refm, _ = make_ref.load()
    boards = synthetic_packed(400, seed=77, max_exp=5)
    tiles = oracle.unpack(boards).astype(np.int64)
    for t in tiles:
        assert tuple(refm.legal_actions_mask(t)) == npe.legal_actions_mask(t)
        assert ref.is_done(t) == npe.is_done(t)
        for a in range(4):
            s1, r1 = ref.latent_state(t, a)
            s2, r2 = npe.latent_state(t, a)
            np.testing.assert_array_equal(s1, s2)
            assert r1 == r2
        c1, c2 = ref.after_state(t), npe.after_state(t)
        assert len(c1) == len(c2)
        for (x, p), (y, q) in zip(c1, c2):
            np.testing.assert_array_equal(x, y)
            assert p == q
    # same generator state => same spawns
    r1, r2 = np.random.default_rng(5), np.random.default_rng(5)
    ref._GENERATOR = r1
    for t in tiles[:100]:
        a, ra = ref.next_state(t.copy(), 1)
        b, rb = npe.next_state(t.copy(), 1, r2)
        np.testing.assert_array_equal(a, b)
        assert ra == rb


# ---------------------------------------------------------------- presentation helpers (utils/)

def test_oracle_encodings_match_reference(golden):
    """utils/binary.py encode_flatten(., 31) and utils/normalize.py normalize_reward."""
    g = golden('encodings')
    for tiles, want in zip(g['tiles'], g['onehot']):
        np.testing.assert_array_equal(npe.encode_flatten(tiles.astype(np.int64), 31), want)
    for r, want in zip(g['rewards'], g['normalized']):
        assert npe.normalize_reward(float(r)) == want
    assert g['normalized'][0] == 0.0 and g['normalized'][1] == 2 / 16


# ---------------------------------------------------------------- n-step returns (self_play.py:524)

def test_oracle_nstep_returns_match_reference(golden):
    """compute_n_step_returns of the real reference on 7 trajectories (T = 1 .. 40; lambda 0, 1;
    n_steps 1; discount 1).  Same arithmetic order, so the restatement agrees to rounding."""
    g = golden('nstep')
    for i, (disc, n_steps, lam) in enumerate(g['params']):
        got = npe.n_step_returns(g[f'r{i}'].astype(np.float64), g[f'v{i}'].astype(np.float64), disc, int(n_steps), lam)
        np.testing.assert_allclose(got, g[f'out{i}'], rtol=1e-13, atol=0)


# ---------------------------------------------------------------- expectimax backup (row f2)

def test_oracle_expectimax_matches_reference(golden):
    """Depth-1 expectimax values computed with the real reference's latent_state / after_state."""
    g = golden('expectimax')

    def leaf(tiles):
        t = np.asarray(tiles)
        mx = int(t.max())
        return float((t == 0).sum()) + 0.25 * (mx.bit_length() - 1 if mx else 0)
    for tiles, want in zip(g['tiles'][:96], g['q'][:96]):
        got = npe.expectimax_q(tiles.astype(np.int64), leaf, float(g['discount']))
        assert np.array_equal(np.isinf(got), np.isinf(want))
        np.testing.assert_allclose(got[np.isfinite(want)], want[np.isfinite(want)], rtol=1e-12)


def test_oracle_expectimax_depth2_matches_reference(golden):
    g = golden('expectimax')

    def leaf(tiles):
        t = np.asarray(tiles)
        mx = int(t.max())
        return float((t == 0).sum()) + 0.25 * (mx.bit_length() - 1 if mx else 0)
    for tiles, want in zip(g['deep_tiles'][:4], g['deep_q'][:4]):
        got = npe.expectimax_q(tiles.astype(np.int64), leaf, float(g['discount']), depth=2)
        assert np.array_equal(np.isinf(got), np.isinf(want))
        np.testing.assert_allclose(got[np.isfinite(want)], want[np.isfinite(want)], rtol=1e-12)
