"""NumPy restatement of the two reference engines — TEST INFRASTRUCTURE ONLY.

Part 1 restates the NumPy engine (``twentyfortyeight/core/gameboard.py``, ``gamemove.py``,
``envs/twentyfortyeight.py``) one board per call, with the same kind of NumPy work per step as
the reference (rot90 views, a Python merge loop per row, argwhere + Generator.choice spawns), so
that timing it on the GPU box's host cores is representative of the reference engine.  It is
the ``cpu_baseline`` of kind "port" in bench.py.

Part 2 restates the semantics of the JAX engine (``reinforce/game/core.py``, ``env.py``) in
NumPy.  JAX is not installable in the build image, so this is a *restatement*, never "JAX":
it is pinned only by the literal vectors in the reference's tests/test_jax_game.py and by
agreement with part 1 where the two engines coincide.

Parity status: PINNED for part 1 (tests/test_oracle_golden.py compares it with fixtures the
real reference produced, and — in the build container — with the imported reference itself).
Citations are relative to the reference repository root.
"""

from __future__ import annotations

import numpy as np

P_TWO, P_FOUR = 0.9, 0.1  # gameboard.py:13


# =====================================================================================
# Part 1 — NumPy engine restatement
# =====================================================================================

def merge_line(line: np.ndarray):
    """gameboard.py:23-69 — drop zeros, pair equal neighbours greedily from index 0, each tile
    merging at most once.  Returns (score, trimmed array)."""
    tiles = line[line != 0]
    count = len(tiles)
    if count <= 1:
        return 0, tiles
    kept, gained, pos = [], 0, 0
    while pos < count - 1:
        if tiles[pos] == tiles[pos + 1]:
            double = tiles[pos] * 2
            kept.append(double)
            gained += double
            pos += 2
        else:
            kept.append(tiles[pos])
            pos += 1
    if pos == count - 1:
        kept.append(tiles[-1])
    return gained, np.array(kept, dtype=line.dtype)


def slide_left(board: np.ndarray):
    """gameboard.py:72-102 — every row through merge_line, zero padded on the right.
    Returns (score as float, board)."""
    out = np.zeros_like(board)
    total = 0.0
    for r, row in enumerate(board):
        gained, merged = merge_line(row)
        total += gained
        out[r, : len(merged)] = merged
    return total, out


def latent_state(board: np.ndarray, action: int):
    """gameboard.py:105-129 — rotate by `action` quarter turns, slide left, rotate back."""
    turned = np.rot90(board, k=action)
    reward, moved = slide_left(turned)
    return np.rot90(moved, k=-action), reward


def can_move_left(board: np.ndarray) -> bool:
    """gamemove.py:132-164."""
    left, right = board[:, :-1], board[:, 1:]
    if ((left == 0) & (right != 0)).any():
        return True
    return bool(((left != 0) & (left == right)).any())


def legal_actions_mask(board: np.ndarray):
    """gamemove.py:45-83 — (left, up, right, down)."""
    left, right = board[:, :-1], board[:, 1:]
    top, bottom = board[:-1, :], board[1:, :]
    h_merge = ((left != 0) & (left == right)).any()
    v_merge = ((top != 0) & (top == bottom)).any()
    return (
        bool(((left == 0) & (right != 0)).any() or h_merge),
        bool(((top == 0) & (bottom != 0)).any() or v_merge),
        bool(((right == 0) & (left != 0)).any() or h_merge),
        bool(((bottom == 0) & (top != 0)).any() or v_merge),
    )


def legal_actions(board: np.ndarray):
    """gamemove.py:109."""
    mask = legal_actions_mask(board)
    return [a for a in range(4) if mask[a]]


def fill_cells(board: np.ndarray, count: int, rng: np.random.Generator) -> np.ndarray:
    """gameboard.py:247-288 — values drawn first, then cell ranks without replacement; in place."""
    free = np.argwhere(board == 0)
    if len(free) == 0:
        return board
    count = min(count, len(free))
    values = rng.choice([2, 4], size=count, p=[P_TWO, P_FOUR])
    picks = rng.choice(len(free), size=count, replace=False)
    board[tuple(free[picks].T)] = values
    return board


def next_state(board: np.ndarray, action: int, rng: np.random.Generator):
    """gameboard.py:291-325 — legal: slide, unrotate, one spawn; illegal: (board, 0)."""
    turned = np.rot90(board, k=action)
    if can_move_left(turned):
        reward, moved = slide_left(turned)
        board = np.rot90(moved, k=-action)
        return fill_cells(board, 1, rng), reward
    return board, 0


def is_done(board: np.ndarray) -> bool:
    """gameboard.py:328-348."""
    return bool(
        np.all(board != 0)
        and not np.any(board[:-1] == board[1:])
        and not np.any(board[:, :-1] == board[:, 1:])
    )


def after_state(board: np.ndarray):
    """gameboard.py:132-171."""
    free = np.argwhere(board == 0)
    if len(free) == 0:
        return [(board, 1.0)]
    out = []
    for cell in free:
        for value, p in ((2, P_TWO), (4, P_FOUR)):
            child = board.copy()
            child[tuple(cell)] = value
            out.append((child, p / len(free)))
    return out


def encode_flatten(board: np.ndarray, width: int = 31) -> np.ndarray:
    """utils/binary.py:11-86 — one-hot of log2(tile) per cell (an empty cell is index 0), flattened."""
    flat = board.reshape(-1)
    exps = np.zeros(flat.size, dtype=np.int64)
    nz = flat != 0
    exps[nz] = np.log2(flat[nz].astype(np.float64)).astype(np.int64)
    out = np.zeros((flat.size, width), dtype=np.int64)
    out[np.arange(flat.size), exps] = 1
    return out.reshape(-1)


def normalize_reward(reward: float, max_tile: int = 2 ** 16) -> float:
    """utils/normalize.py:6-33."""
    return 0.0 if reward == 0 else float(np.log2(reward) / np.log2(max_tile))


def expectimax_q(board: np.ndarray, value_fn, discount: float = 1.0, depth: int = 1,
                 terminal_value: float = 0.0) -> np.ndarray:
    """Depth-`depth` expectimax action values with the reference's own building blocks
    (gameboard.py:105 latent_state, :132 after_state incl. its probabilities, gamemove.py:45):
    q[a] = reward + discount * sum_i p_i * V(child_i), V = value_fn at the leaves and the max over
    legal actions (terminal_value when there is none) above; -inf for an illegal action."""
    def child_value(child):
        if depth == 1:
            return value_fn(child)
        q_c = expectimax_q(np.asarray(child), value_fn, discount, depth - 1, terminal_value)
        best = q_c.max()
        return terminal_value if np.isinf(best) else best

    q = np.full(4, -np.inf)
    legal = legal_actions_mask(board)
    for a in range(4):
        if not legal[a]:
            continue
        after, reward = latent_state(board, a)
        q[a] = reward + discount * sum(p * child_value(child) for child, p in after_state(np.array(after)))
    return q


def n_step_returns(rewards: np.ndarray, values: np.ndarray, discount: float, n_steps: int,
                   td_lambda: float = 0.5) -> np.ndarray:
    """reinforce/training/self_play.py:524-579 — TD(lambda)-weighted n-step returns of ONE
    trajectory (float64): G_n = sum_{i<n} g^i r[t+i] (+ g^n v[t+n] if t+n < T), weights
    (1-l) l^(n-1) for n < n_steps and l^(n-1) for n = n_steps, normalised by their sum."""
    T = len(rewards)
    out = np.zeros(T)
    for t in range(T):
        total = weight_sum = 0.0
        for n in range(1, min(n_steps, T - t) + 1):
            g_n = float(sum(discount ** i * rewards[t + i] for i in range(n)))
            if t + n < T:
                g_n += discount ** n * values[t + n]
            w = (1 - td_lambda) * td_lambda ** (n - 1) if n < n_steps else td_lambda ** (n - 1)
            total += w * g_n
            weight_sum += w
        out[t] = total / weight_sum if weight_sum > 0 else values[t]
    return out


class Game:
    """twentyfortyeight.py:10-140 (size 4, raw observation, raw reward)."""

    def __init__(self, rng: np.random.Generator):
        self.rng = rng
        self.board = np.zeros((4, 4), dtype=np.int64)
        self.reward = 0.0

    def reset(self, seed=None) -> np.ndarray:
        self.board = np.zeros((4, 4), dtype=np.int64)
        rng = np.random.default_rng(seed) if seed is not None else self.rng  # gameboard.py:272
        self.board = fill_cells(self.board, 2, rng)
        self.reward = 0
        return self.board

    def step(self, action: int):
        self.board, self.reward = next_state(self.board, action, self.rng)
        return self.board, self.reward, is_done(self.board)


def play_random_games(first_game: int, n_games: int, policy_seed: int = 0):
    """BASELINE.json config 1 / SURVEY.md §8d: game g starts from reset(seed=g); uniform-legal
    policy from default_rng(policy_seed); spawns from default_rng(PCG64DXSM(0)); play to
    game-over.  Returns (total steps, total score, list of max tiles)."""
    rng = np.random.default_rng(np.random.PCG64DXSM(0))
    policy = np.random.default_rng(policy_seed)
    steps = score = 0
    tops = []
    for g in range(first_game, first_game + n_games):
        game = Game(rng)
        game.reset(seed=g)
        while True:
            legal = legal_actions(game.board)
            if not legal:
                break
            _, r, done = game.step(legal[policy.integers(0, len(legal))])
            steps += 1
            score += int(r)
            if done:
                break
        tops.append(int(game.board.max()))
    return steps, score, tops


def step_boards(tiles: np.ndarray, actions: np.ndarray, seed: int = 0):
    """Config-2 workload through the Python engine: next_state per board (one Python call per
    env, as a user of the reference would run it).  Returns (boards, rewards)."""
    rng = np.random.default_rng(seed)
    out = np.empty_like(tiles)
    rewards = np.empty(len(tiles), dtype=np.float64)
    for i in range(len(tiles)):
        out[i], rewards[i] = next_state(tiles[i], int(actions[i]), rng)
    return out, rewards


# =====================================================================================
# Part 2 — JAX engine semantics, restated in NumPy (reinforce/game/core.py, env.py)
# =====================================================================================

def jx_merge_row(row: np.ndarray):
    """core.py:43-101 — compact, three-step merge scan with the 'previous merged' carry,
    compact again.  Returns (row[4], score)."""
    def compact(r):
        nz = r[r != 0]
        return np.concatenate([nz, np.zeros(4 - len(nz), dtype=r.dtype)])

    arr = compact(np.asarray(row).copy())
    score, prev = 0, False
    for i in range(3):  # core.py:92-96 scan over pairs (0,1), (1,2), (2,3)
        can = arr[i] == arr[i + 1] and arr[i] != 0 and not prev
        if can:
            arr[i] *= 2
            arr[i + 1] = 0
            score += int(arr[i])
        prev = bool(can)
    return compact(arr), score


def jx_latent_state(board: np.ndarray, action: int):
    """core.py:129-170 — (board, reward): board FIRST."""
    turned = np.rot90(board, k=action)
    rows, score = [], 0
    for r in turned:
        m, s = jx_merge_row(r)
        rows.append(m)
        score += s
    return np.rot90(np.stack(rows), k=-action), score


def jx_legal_actions_mask(board: np.ndarray) -> np.ndarray:
    """core.py:275-302 — action legal iff it changes the board."""
    return np.array([bool(np.any(jx_latent_state(board, a)[0] != board)) for a in range(4)])


def jx_is_done(board: np.ndarray) -> bool:
    """core.py:305-321 — no legal action (true for the empty board as well)."""
    return not jx_legal_actions_mask(board).any()


def jx_fill_cell(board: np.ndarray, rank: int, value: int) -> np.ndarray:
    """core.py:173-231 with the two random draws injected: `rank`-th empty cell in row-major
    order (cumsum/argmax, :206-213) receives `value`; a full board is returned unchanged."""
    flat = board.reshape(-1).copy()
    free = np.flatnonzero(flat == 0)
    if len(free) == 0:
        return board
    flat[free[rank]] = value
    return flat.reshape(4, 4)


def jx_next_state(board: np.ndarray, action: int, rank: int, value: int):
    """core.py:234-272 — spawn only if the board changed; reward float32, 0 if unchanged."""
    after, reward = jx_latent_state(board, action)
    changed = bool(np.any(after != board))
    if changed:
        after = jx_fill_cell(after, rank, value)
    return after, np.float32(reward if changed else 0.0)


def jx_env_step(board, step_count, done, total_reward, action, rank, value):
    """env.py:79-133 — done envs pass through with reward 0 and done True; otherwise step_count
    increments even for an illegal action and total_reward accumulates in float32."""
    if done:
        return (board, step_count, True, total_reward), np.float32(0.0), True
    new_board, reward = jx_next_state(board, action, rank, value)
    game_done = jx_is_done(new_board)
    state = (new_board, step_count + 1, game_done, np.float32(total_reward + reward))
    return state, reward, game_done


def jx_encode_observation(board: np.ndarray) -> np.ndarray:
    """core.py:346-377."""
    b = board.astype(np.float32)
    logs = np.where(board > 0, np.log2(np.where(board > 0, b, 1.0)), 0.0).astype(np.float32)
    return (logs / np.float32(16.0)).reshape(-1)


def jx_action_distribution(temperature: float, policy: np.ndarray, legal: np.ndarray) -> np.ndarray:
    """core.py:401-421 — the distribution sample_action draws from (float32 arithmetic)."""
    pol = np.where(legal, policy.astype(np.float32), np.float32(0.0))
    total = pol.sum(dtype=np.float32)
    n_legal = np.float32(legal.sum())
    uniform = legal.astype(np.float32) / max(n_legal, np.float32(1.0))
    pol = uniform if total < 1e-8 else pol / total
    if temperature < 0.01:
        out = np.zeros(4, dtype=np.float32)
        out[int(np.argmax(pol))] = 1.0
        return out
    z = np.log(pol + np.float32(1e-8)) / np.float32(temperature)
    z = np.exp(z - z.max())
    return (z / z.sum()).astype(np.float32)
