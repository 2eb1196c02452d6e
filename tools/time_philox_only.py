"""Philox-mode g2048_step on 2^26 boards at 512 / 640 / 768 threads per CTA (CUDA events, 20 launches)."""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
from bench import synthetic_packed
from simulate_2048_b200 import ops
ops.init()
n = 1 << 26
b = torch.from_numpy(np.tile(synthetic_packed(1 << 20), n >> 20).view(np.int64)).cuda()
a = torch.randint(0, 4, (n,), dtype=torch.uint8, device='cuda')
out = ops.step(b, a, None, seed=1)
for threads in (512, 640, 768):
    ops.set_tuning(threads, 4)
    for _ in range(3):
        ops.step(b, a, None, seed=1, out=out)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for i in range(20):
        ops.step(b, a, None, seed=1, step_base=i, out=out)
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / 20
    print(threads, 'philox', round(ms, 4), 'ms', round(n / ms / 1e6, 1), 'G env-steps/s', round(n * 22 / ms / 1e6 / 6543.4, 4))
