"""Device time per g2048_expand / g2048_expand_compact call at small batch sizes: CUDA graph of 40
back-to-back calls into preallocated buffers (as bench.py's secondary.expand)."""
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
from bench import synthetic_packed  # noqa: E402
from simulate_2048_b200 import ops  # noqa: E402

ops.init()
base = synthetic_packed(1 << 20)
for n in (4096, 16384, 65536, 131071):
    b = torch.from_numpy(base[:n].view(np.int64).copy()).cuda()
    res = {}
    for name in ('compact_scan', 'compact_given_offsets', 'dense'):
        if name == 'compact_scan':
            ex = ops.expand(b, children='compact', capacity=72 * n)
            fn = lambda: ops.expand(b, children='compact', out=ex)  # noqa: E731
        elif name == 'compact_given_offsets':
            ex = ops.expand(b, children='compact')
            offs = ex.child_offsets.clone()
            fn = lambda: ops.expand(b, children='compact', out=ex, child_offsets=offs)  # noqa: E731
        else:
            ex = ops.expand(b, children='dense')
            fn = lambda: ops.expand(b, children='dense', out=ex)  # noqa: E731
        fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):
            with torch.cuda.graph(g, stream=side):
                for _ in range(40):
                    fn()
        torch.cuda.synchronize()
        for _ in range(3):
            g.replay()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        for _ in range(5):
            g.replay()
        ev[1].record()
        torch.cuda.synchronize()
        res[name] = round(ev[0].elapsed_time(ev[1]) * 1000 / 200, 2)
    print(n, res, 'us per call')
