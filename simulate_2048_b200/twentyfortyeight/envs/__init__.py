"""``TwentyFortyEight`` environment class — mirrors ``twentyfortyeight/envs`` of the reference."""

from .twentyfortyeight import TwentyFortyEight  # noqa: E402,F401  (manuals_control.py:7 imports it from here)

__all__ = ['TwentyFortyEight']
