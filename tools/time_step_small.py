"""Device time of g2048_step at small and medium batch sizes (CUDA graph of 20 launches)."""
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
from bench import synthetic_packed  # noqa: E402
from simulate_2048_b200 import ops  # noqa: E402

ops.init()
base = torch.from_numpy(synthetic_packed(1 << 20).view(np.int64)).cuda()
for impl in (0, 1):
    ops.set_impl(impl)
    for lg in (12, 14, 16, 17, 18, 19, 20):
        n = 1 << lg
        b = base[:n].contiguous()
        a = torch.randint(0, 4, (n,), dtype=torch.uint8, device='cuda')
        out = ops.step(b, a, None, seed=1)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):
            with torch.cuda.graph(g, stream=side):
                for _ in range(20):
                    ops.step(b, a, None, seed=1, out=out)
        torch.cuda.synchronize()
        g.replay()
        torch.cuda.synchronize()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        e[0].record()
        for _ in range(5):
            g.replay()
        e[1].record()
        torch.cuda.synchronize()
        us = e[0].elapsed_time(e[1]) * 1000 / 100
        print('impl', impl, 'n', n, round(us, 2), 'us/launch', round(n / us / 1e3, 2), 'G env-steps/s')
ops.set_impl(0)
