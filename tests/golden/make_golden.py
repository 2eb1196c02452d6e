"""Generate the golden fixtures in tests/golden/ by running the REAL reference NumPy engine.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_golden.py

It imports ``twentyfortyeight`` from /root/reference (read-only), never copies its source,
and writes small compressed .npz files that the oracle and the CUDA path are checked
against.  Boards are stored both as int64 tile values and as packed u64 words
(nibble i = log2(tile) of flat cell i = 4r + c; see include/g2048.h).

Fixtures
  rows.npz        whole row table from merge_column (gameboard.py:23) on values 2**e
  synthetic.npz   N=4096 random boards (generator of tests/test_perf_utils.py:12-19,
                  vectorised as SURVEY.md §8d, seed 2048): latent_state x4, masks, is_done,
                  after_state children, next_state(seed=i) with the spawn recovered by diff
  games.npz       48 full games of TwentyFortyEight (reset(seed=g), uniform-legal policy):
                  every (board, action, next board, reward, done)
  fill_cells.npz  fill_cells draws on a fixed board, seeds 0..63, and reset(seed) boards
  encodings.npz   utils/binary.py encode_flatten(., 31) of 64 boards; utils/normalize.py values
  nstep.npz       reinforce/training/self_play.py compute_n_step_returns on 7 random trajectories
  expectimax.npz  depth-1 expectimax values of 256 boards: latent_state + after_state probabilities
"""

from __future__ import annotations

import hashlib
import os
import sys

import numpy as np

sys.path.insert(0, '/root/reference')

from twentyfortyeight.core import gameboard  # noqa: E402
from twentyfortyeight.core.gameboard import (  # noqa: E402
    after_state,
    fill_cells,
    is_done,
    latent_state,
    merge_column,
    next_state,
)
from twentyfortyeight.core.gamemove import legal_actions, legal_actions_mask  # noqa: E402
from twentyfortyeight.envs.twentyfortyeight import TwentyFortyEight  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def pack(tiles: np.ndarray) -> np.uint64:
    """int tile values [4,4] -> packed u64 (exponent > 15 wraps, flagged by callers)."""
    flat = np.asarray(tiles, dtype=np.int64).reshape(16)
    word = 0
    for i, t in enumerate(flat.tolist()):
        e = int(t).bit_length() - 1 if t else 0
        word |= (e & 0xF) << (4 * i)
    return np.uint64(word)


def synthetic_boards(n: int, seed: int = 2048) -> np.ndarray:
    """SURVEY.md §8d recipe: n_tiles ~ U{1..16}, positions = first n of a random order,
    exponents ~ U{1..11}; draws in the order integers, random, integers."""
    rng = np.random.default_rng(seed)
    cnt = rng.integers(1, 17, n)
    keys = rng.random((n, 16))
    ex = rng.integers(1, 12, (n, 16))
    order = np.argsort(keys, axis=1, kind='stable')
    exps = np.zeros((n, 16), dtype=np.int64)
    for i in range(n):
        exps[i, order[i, : cnt[i]]] = ex[i, : cnt[i]]
    return np.where(exps > 0, 1 << exps, 0).astype(np.int64).reshape(n, 4, 4)


def make_rows() -> None:
    result = np.zeros(65536, dtype=np.uint16)
    score = np.zeros(65536, dtype=np.uint32)
    for r in range(65536):
        col = np.array([(1 << ((r >> (4 * j)) & 0xF)) if (r >> (4 * j)) & 0xF else 0 for j in range(4)],
                       dtype=np.int64)
        s, merged = merge_column(col)
        res = 0
        for j, t in enumerate(merged.tolist()):
            res |= ((int(t).bit_length() - 1) & 0xF) << (4 * j)
        result[r] = res
        score[r] = s
    np.savez_compressed(os.path.join(HERE, 'rows.npz'), result=result, score=score)
    print('rows', hashlib.sha256(result.tobytes()).hexdigest()[:16], int(score.sum()))


def make_synthetic(n: int = 4096) -> None:
    boards = synthetic_boards(n)
    packed = np.array([pack(b) for b in boards], dtype=np.uint64)
    after = np.zeros((n, 4), dtype=np.uint64)
    rewards = np.zeros((n, 4), dtype=np.int64)
    mask = np.zeros(n, dtype=np.uint8)
    done = np.zeros(n, dtype=np.uint8)
    child_count = np.zeros((n, 4), dtype=np.uint8)
    children = []
    child_probs = []
    rng = np.random.default_rng(7)
    actions = rng.integers(0, 4, n).astype(np.uint8)
    nxt = np.zeros(n, dtype=np.uint64)
    nxt_reward = np.zeros(n, dtype=np.int64)
    nxt_spawn = np.zeros(n, dtype=np.uint8)  # ABI spawn byte recovered by diff (0xFF = no spawn)
    for i in range(n):
        b = boards[i]
        m = legal_actions_mask(b)
        mask[i] = sum(int(v) << a for a, v in enumerate(m))
        done[i] = int(is_done(b))
        for a in range(4):
            s, r = latent_state(b, a)
            after[i, a] = pack(s)
            rewards[i, a] = int(r)
            if m[a]:
                outs = after_state(np.array(s))
                child_count[i, a] = len(outs)
                for st, p in outs:
                    children.append(pack(st))
                    child_probs.append(p)
        a = int(actions[i])
        s_after, _ = latent_state(b, a)
        nb, r = next_state(b.copy(), a, seed=i)
        nxt[i] = pack(nb)
        nxt_reward[i] = int(r)
        diff = np.argwhere(np.asarray(nb).reshape(16) != np.asarray(s_after).reshape(16)).reshape(-1)
        if m[a]:
            assert diff.size == 1
            cell = int(diff[0])
            val = int(np.asarray(nb).reshape(16)[cell])
            assert val in (2, 4)
            empties = np.argwhere(np.asarray(s_after).reshape(16) == 0).reshape(-1).tolist()
            nxt_spawn[i] = empties.index(cell) | (0x10 if val == 4 else 0)
        else:
            assert diff.size == 0 and r == 0
            nxt_spawn[i] = 0xFF
    np.savez_compressed(
        os.path.join(HERE, 'synthetic.npz'),
        tiles=boards.astype(np.int32), packed=packed, after=after, rewards=rewards, mask=mask,
        done=done, child_count=child_count, children=np.array(children, dtype=np.uint64),
        child_probs=np.array(child_probs, dtype=np.float64), actions=actions, next=nxt,
        next_reward=nxt_reward, next_spawn=nxt_spawn,
    )
    print('synthetic packed', hashlib.sha256(packed.tobytes()).hexdigest()[:16],
          'after', hashlib.sha256(after.tobytes()).hexdigest()[:16],
          'rewards sum', int(rewards.sum()), 'children', len(children))


def make_games(n_games: int = 48) -> None:
    gameboard._GENERATOR = np.random.default_rng(np.random.PCG64DXSM(0))  # deterministic spawns
    policy = np.random.default_rng(0)
    rec = {k: [] for k in ('game', 'board', 'action', 'next', 'reward', 'done', 'spawn')}
    first = []
    for g in range(n_games):
        env = TwentyFortyEight()
        env.reset(seed=g)
        first.append(pack(env._current_state))
        while True:
            legal = legal_actions(env._current_state)
            if not legal:
                break
            # mix in an occasional illegal action so the no-op path is pinned too
            if policy.random() < 0.05:
                a = int(policy.integers(0, 4))
            else:
                a = int(legal[policy.integers(0, len(legal))])
            before = env._current_state.copy()
            s_after, _ = latent_state(before, a)
            obs, r, d = env.step(a)
            diff = np.argwhere(obs.reshape(16) != np.asarray(s_after).reshape(16)).reshape(-1)
            if diff.size == 1:
                cell = int(diff[0])
                empties = np.argwhere(np.asarray(s_after).reshape(16) == 0).reshape(-1).tolist()
                sp = empties.index(cell) | (0x10 if int(obs.reshape(16)[cell]) == 4 else 0)
            else:
                sp = 0xFF
            rec['game'].append(g)
            rec['board'].append(pack(before))
            rec['action'].append(a)
            rec['next'].append(pack(obs))
            rec['reward'].append(int(r))
            rec['done'].append(int(d))
            rec['spawn'].append(sp)
            if d:
                break
    np.savez_compressed(
        os.path.join(HERE, 'games.npz'),
        first=np.array(first, dtype=np.uint64), game=np.array(rec['game'], dtype=np.int32),
        board=np.array(rec['board'], dtype=np.uint64), action=np.array(rec['action'], dtype=np.uint8),
        next=np.array(rec['next'], dtype=np.uint64), reward=np.array(rec['reward'], dtype=np.int64),
        done=np.array(rec['done'], dtype=np.uint8), spawn=np.array(rec['spawn'], dtype=np.uint8),
    )
    print('games', n_games, 'steps', len(rec['game']), 'mean len', len(rec['game']) / n_games)


def make_fill_cells() -> None:
    z = np.array([[2, 0, 0, 0], [0, 0, 4, 0], [0, 0, 0, 0], [0, 8, 0, 0]], dtype=np.int64)
    out = []
    for seed in range(64):
        b = fill_cells(z.copy(), 1, seed)
        cell = int(np.argwhere(b.reshape(16) != z.reshape(16))[0][0])
        empties = np.argwhere(z.reshape(16) == 0).reshape(-1).tolist()
        out.append((cell, int(b.reshape(16)[cell]), empties.index(cell)))
    resets = np.array([pack(TwentyFortyEight().reset(seed=s)) for s in range(64)], dtype=np.uint64)
    np.savez_compressed(os.path.join(HERE, 'fill_cells.npz'), base=pack(z),
                        draws=np.array(out, dtype=np.int32), resets=resets,
                        numpy_version=np.array(np.__version__))
    print('fill_cells', out[:8], [hex(int(x)) for x in resets[:2]])


def make_encodings() -> None:
    from twentyfortyeight.utils.binary import encode_flatten
    from twentyfortyeight.utils.normalize import normalize_reward
    boards = synthetic_boards(64, seed=31)
    onehot = np.stack([encode_flatten(b, encodage_size=31) for b in boards]).astype(np.uint8)
    rewards = np.array([0, 4, 8, 12, 20, 60, 84, 2048, 4100, 65536, 131072], dtype=np.int64)
    norm = np.array([normalize_reward(float(r)) for r in rewards], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, 'encodings.npz'), tiles=boards.astype(np.int32), onehot=onehot,
                        rewards=rewards, normalized=norm)
    print('encodings', onehot.shape, norm[:4])


def leaf_value(tiles) -> float:
    """Deterministic stand-in for a value network: empty cells + log2(max tile) / 4 (exact in f32)."""
    t = np.asarray(tiles)
    mx = int(t.max())
    return float((t == 0).sum()) + 0.25 * (mx.bit_length() - 1 if mx else 0)


def make_expectimax(n: int = 256) -> None:
    """Depth-1 expectimax values from the real reference: latent_state + after_state (with its
    probabilities) + legal_actions_mask, leaf values from `leaf_value`."""
    boards = synthetic_boards(n, seed=77)
    q = np.full((n, 4), -np.inf)
    for i, b in enumerate(boards):
        legal = legal_actions_mask(b)
        for a in range(4):
            if legal[a]:
                s_after, r = latent_state(b, a)
                q[i, a] = r + 0.99 * sum(p * leaf_value(child) for child, p in after_state(np.array(s_after)))
    # depth 2 on a few crowded boards (few empty cells keep the tree small): the children's value is
    # the max over their legal actions of the depth-1 value, 0 for a dead child
    def q_depth1(b):
        out = np.full(4, -np.inf)
        legal = legal_actions_mask(b)
        for a in range(4):
            if legal[a]:
                s_after, r = latent_state(b, a)
                out[a] = r + 0.99 * sum(p * leaf_value(child) for child, p in after_state(np.array(s_after)))
        return out
    deep = [b for b in synthetic_boards(400, seed=78) if (b == 0).sum() <= 5][:12]
    q2 = np.full((len(deep), 4), -np.inf)
    for i, b in enumerate(deep):
        legal = legal_actions_mask(b)
        for a in range(4):
            if legal[a]:
                s_after, r = latent_state(b, a)
                total = 0.0
                for child, p in after_state(np.array(s_after)):
                    best = q_depth1(np.asarray(child)).max()
                    total += p * (0.0 if np.isinf(best) else best)
                q2[i, a] = r + 0.99 * total
    np.savez_compressed(os.path.join(HERE, 'expectimax.npz'), tiles=boards.astype(np.int32),
                        packed=np.array([pack(b) for b in boards], dtype=np.uint64), q=q, discount=np.array(0.99),
                        deep_tiles=np.array(deep, dtype=np.int32),
                        deep_packed=np.array([pack(b) for b in deep], dtype=np.uint64), deep_q=q2)
    print('expectimax', n, q[:2])


def make_nstep() -> None:
    """compute_n_step_returns lives in a module that imports jax (absent here): the function is
    plain NumPy, so its definition alone is compiled from the reference source and executed."""
    import ast
    path = '/root/reference/reinforce/training/self_play.py'
    tree = ast.parse(open(path).read())
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == 'compute_n_step_returns'][0]
    ns = {'np': np}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), path, 'exec'), ns)
    ref = ns['compute_n_step_returns']
    rng = np.random.default_rng(524)
    cases = []
    for T, n_steps, disc, lam in ((1, 10, 0.999, 0.5), (7, 10, 0.999, 0.5), (40, 10, 0.999, 0.5),
                                  (40, 5, 0.97, 0.8), (33, 1, 0.9, 0.5), (25, 4, 1.0, 0.0), (25, 4, 0.95, 1.0)):
        r = (rng.integers(0, 64, T) * 4).astype(np.float32)
        v = rng.normal(200.0, 80.0, T).astype(np.float32)
        cases.append((r, v, disc, n_steps, lam, ref(r.astype(np.float64), v.astype(np.float64), disc, n_steps, lam)))
    np.savez_compressed(
        os.path.join(HERE, 'nstep.npz'),
        **{f'r{i}': c[0] for i, c in enumerate(cases)}, **{f'v{i}': c[1] for i, c in enumerate(cases)},
        **{f'out{i}': c[5] for i, c in enumerate(cases)},
        params=np.array([[c[2], c[3], c[4]] for c in cases], dtype=np.float64))
    print('nstep', len(cases), cases[2][5][:3])


if __name__ == '__main__':
    make_expectimax()
    make_nstep()
    make_encodings()
    make_rows()
    make_synthetic()
    make_games()
    make_fill_cells()
