"""Drop-in for the NumPy engine ``twentyfortyeight`` of simulate_2048, executed on B200.

Same modules, names, argument and return conventions as the reference
(``twentyfortyeight/core/gameboard.py``, ``core/gamemove.py``, ``envs/twentyfortyeight.py``);
boards are ``numpy`` int arrays of tile values, every game-rule computation runs in the CUDA
kernels behind the C ABI (no CPU implementation of the rules exists in this package).
"""
