"""Workloads of BASELINE.json run on the REAL reference engine staged under ``oracle/_ref`` (see
``make_ref.py``) — test / bench infrastructure, never imported by the product.

Each function calls the reference's own public functions once per board, as a user of the
reference does (it has no batched NumPy API): ``gameboard.next_state`` (gameboard.py:291),
``gamemove.legal_actions_mask`` (gamemove.py:45), ``gameboard.latent_state`` (:105),
``gameboard.after_state`` (:132) and ``TwentyFortyEight.reset / step``
(envs/twentyfortyeight.py:84-133).  The module generator ``gameboard._GENERATOR`` is unseeded in
the reference (gameboard.py:20); it is rebound to a seeded PCG64DXSM generator so runs repeat.
"""

from __future__ import annotations

import numpy as np
from numpy.random import PCG64DXSM, default_rng

from . import make_ref


def step_boards(tiles: np.ndarray, actions: np.ndarray, seed: int = 0):
    """Config 2 on the reference: one ``next_state(state, action)`` call per board.
    tiles int64[n,4,4].  Returns (new tiles, rewards)."""
    gb, _, _ = make_ref.load()
    gb._GENERATOR = default_rng(PCG64DXSM(seed))
    out = np.empty_like(tiles)
    rewards = np.empty(len(tiles), dtype=np.float64)
    for i in range(len(tiles)):
        out[i], rewards[i] = gb.next_state(tiles[i], int(actions[i]))
    return out, rewards


def expand_boards(tiles: np.ndarray) -> int:
    """Config 3 on the reference: legal mask, four afterstates + rewards, every chance child of each
    legal afterstate.  Returns the number of children produced."""
    gb, gm, _ = make_ref.load()
    children = 0
    for b in tiles:
        mask = gm.legal_actions_mask(b)
        for a in range(4):
            after, _ = gb.latent_state(b, a)
            if mask[a]:
                children += len(gb.after_state(np.array(after)))
    return children


def play_random_games(first_game: int, count: int):
    """Config 1 (SURVEY.md §8d / BASELINE.md §4.1): game g = ``TwentyFortyEight().reset(seed=g)``,
    uniform policy over ``legal_actions(state)`` from ``default_rng(g)``, until ``is_finished``.
    Returns (total steps, total score)."""
    gb, gm, Env = make_ref.load()
    steps = score = 0
    for g in range(first_game, first_game + count):
        gb._GENERATOR = default_rng(PCG64DXSM(g))
        policy = default_rng(g)
        env = Env()
        env.reset(seed=g)
        while not env.is_finished:
            legal = gm.legal_actions(env._current_state)
            _, reward, _ = env.step(int(legal[policy.integers(len(legal))]))
            score += int(reward)
            steps += 1
    return steps, score
