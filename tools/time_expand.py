"""Times g2048_expand without children (moves kernel only) and with compact / dense children."""
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
from bench import synthetic_packed  # noqa: E402
from simulate_2048_b200 import ops  # noqa: E402

ops.init()
for log2 in (16, 20, 22):
    boards = torch.from_numpy(synthetic_packed(1 << log2).view(np.int64)).cuda()
    res = {}
    for mode in ('none', 'compact', 'dense'):
        ex = ops.expand(boards, children=mode)
        for _ in range(3):
            ops.expand(boards, children=mode, out=ex)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        for _ in range(10):
            ops.expand(boards, children=mode, out=ex)
        ev[1].record()
        torch.cuda.synchronize()
        res[str(mode)] = round(ev[0].elapsed_time(ev[1]) * 100, 1)  # us per launch
    print(1 << log2, res)
