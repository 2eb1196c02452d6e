"""``TwentyFortyEight`` environment class — mirrors ``twentyfortyeight/envs`` of the reference."""

from .twentyfortyeight import TwentyFortyEight  # noqa: E402,F401  (manuals_control.py:7 imports it from here)

from .vector import VectorTwentyFortyEight  # noqa: E402,F401  (N lock-step games, one launch per step)

__all__ = ['TwentyFortyEight', 'VectorTwentyFortyEight']
