"""Generates tests/golden/rollout_stats.json: the 24-word statistics vector (include/g2048.h
g2048_rollout) of BASELINE config 4 — from-reset random rollouts of envs [0, n), seed 2048 — computed
by the CPU oracle (oracle/g2048_oracle.c orc_rollout_mt, value-domain arithmetic).  bench.py compares
the all-reduced statistics of the sharded GPU run with this vector (`stats_equal_oracle`), and
tests/test_gpu_expand_rollout.py does the same for the smaller sizes.

    python tests/golden/make_rollout_stats.py          # ~4 minutes on 8 cores for 2^24 envs
"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import c_oracle  # noqa: E402

SEED = 2048
lib = c_oracle.lib()
lib.orc_set_num_threads(len(os.sched_getaffinity(0)))
out = {'seed': SEED, 'max_steps': 1 << 20, 'from_reset': True, 'philox_stream': 'v3 game stream (block per env and four draws, mixed-radix digits)',
       'stats': {}}
for lg in (16, 20, 21, 24):
    stats = np.zeros(24, dtype=np.uint64)
    t0 = time.time()
    lib.orc_rollout_mt(C.c_uint64(SEED), C.c_uint64(0), C.c_size_t(1 << lg), C.c_uint32(1 << 20),
                       stats.ctypes.data_as(C.c_void_p))
    out['stats'][str(1 << lg)] = [int(x) for x in stats]
    print(lg, round(time.time() - t0, 1), 's', stats[:7].tolist(), flush=True)
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'rollout_stats.json'), 'w') as f:
    json.dump(out, f, indent=1)
