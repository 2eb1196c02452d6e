"""simulate_2048_b200 — B200-native batched 2048 environment.

The game-transition hot path of schalappe/simulate_2048 (batched reset / step, legal-action
mask, afterstate and chance-spawn expansion, fused random rollouts) as hand-written sm_100a
CUDA behind a C ABI (include/g2048.h), with Python mirrors of the reference's two game APIs:

  simulate_2048_b200.game               <-> reinforce/game        (JAX engine; torch CUDA tensors)
  simulate_2048_b200.twentyfortyeight   <-> twentyfortyeight      (NumPy engine; ndarray in/out)
  simulate_2048_b200.ops                    tensor-level operators over the C ABI
  simulate_2048_b200.parallel               env sharding over the GPUs of one node + statistics

There is no CPU fallback: without the built shared library (``python -m
simulate_2048_b200.build``) every call raises.
"""

__version__ = '0.1.0'
