"""time_step_1m_ab.py [lib.so] — g2048_step (injected spawns) at 2^18 … 2^21 boards per launch, device time per launch
from a CUDA graph of 20 back-to-back launches, with the shipped library or another build of it (A/B)."""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
from simulate_2048_b200 import _cabi
if len(sys.argv) > 1:
    _cabi.LIB_PATH = sys.argv[1]
from bench import synthetic_packed
from simulate_2048_b200 import ops
ops.init()
base = torch.from_numpy(synthetic_packed(1 << 20).view(np.int64)).cuda()
for lg in (18, 19, 20, 21):
    n = 1 << lg
    b = torch.from_numpy(np.tile(synthetic_packed(1 << 20), max(1, n >> 20))[:n].view(np.int64).copy()).cuda()
    a = torch.randint(0, 4, (n,), dtype=torch.uint8, device='cuda')
    sp = torch.randint(0, 2, (n,), dtype=torch.uint8, device='cuda')
    out = ops.step(b, a, sp)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            for _ in range(20):
                ops.step(b, a, sp, out=out)
    torch.cuda.synchronize()
    g.replay(); torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    e[0].record()
    for _ in range(5):
        g.replay()
    e[1].record()
    torch.cuda.synchronize()
    us = e[0].elapsed_time(e[1]) * 1000 / 100
    print('n', n, round(us, 2), 'us/launch', round(n / us / 1e3, 2), 'G env-steps/s', flush=True)
