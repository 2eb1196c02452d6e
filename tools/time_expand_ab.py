"""time_expand_ab.py [lib.so] — compact / dense expansion timings (CUDA events, 20 calls into preallocated
buffers) with the shipped library or, for an A/B, another build of it (path relative to the repo root)."""
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
from simulate_2048_b200 import _cabi  # noqa: E402
if len(sys.argv) > 1:
    _cabi.LIB_PATH = sys.argv[1]
from bench import synthetic_packed  # noqa: E402
from simulate_2048_b200 import ops  # noqa: E402

ops.init()
for log2 in (17, 18, 20, 22):
    boards = torch.from_numpy(synthetic_packed(1 << log2).view(np.int64)).cuda()
    res = {}
    for mode in ('compact', 'dense'):
        ex = ops.expand(boards, children=mode) if mode == 'dense' else ops.expand(boards, children=mode, capacity=72 << log2)
        for _ in range(3):
            ops.expand(boards, children=mode, out=ex)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        for _ in range(20):
            ops.expand(boards, children=mode, out=ex)
        ev[1].record()
        torch.cuda.synchronize()
        res[mode] = round(ev[0].elapsed_time(ev[1]) * 50, 1)  # us per call
    print(1 << log2, res, flush=True)
