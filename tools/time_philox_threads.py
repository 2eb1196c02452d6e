"""Philox-mode and injected-mode g2048_step on 2^26 boards for each CTA size of the smem-table kernel."""
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
from bench import synthetic_packed  # noqa: E402
from simulate_2048_b200 import ops  # noqa: E402

ops.init()
n = 1 << 26
b = torch.from_numpy(np.tile(synthetic_packed(1 << 20), n >> 20).view(np.int64)).cuda()
a = torch.randint(0, 4, (n,), dtype=torch.uint8, device='cuda')
sp = torch.randint(0, 2, (n,), dtype=torch.uint8, device='cuda')
out = ops.step(b, a, None, seed=1)
for threads, envs in ((512, 4), (640, 4), (768, 4), (1024, 4), (512, 8), (768, 8), (1024, 8)):
    ops.set_tuning(threads, envs)
    for name, spawns in (('philox', None), ('injected', sp)):
        for _ in range(3):
            ops.step(b, a, spawns, seed=1, out=out)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        for i in range(20):
            ops.step(b, a, spawns, seed=1, step_base=i, out=out)
        ev[1].record()
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / 20
        byt = 22 if spawns is None else 23
        print(threads, envs, name, round(ms, 4), 'ms', round(n / ms / 1e6, 1), 'G env-steps/s', round(n * byt / ms / 1e6 / 6543.4, 4), 'of measured HBM peak')
ops.set_tuning(0, 4)
