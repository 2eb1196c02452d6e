"""Drop-in for ``twentyfortyeight/utils/normalize.py``.  One scalar per step on the host — a
presentation transform of the reward the device computed, not a game rule."""

from numpy import log2


def normalize_reward(reward: float, max_tile: int = 2 ** (4**2)) -> float:
    """normalize.py:6 — log2(reward) / log2(max_tile), 0 for a zero reward."""
    if reward == 0:
        return 0.0
    return log2(reward) / log2(max_tile)
