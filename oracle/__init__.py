"""CPU oracle for the 2048 transition path — TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import
this package.  The product (``simulate_2048_b200``) never does.
"""
