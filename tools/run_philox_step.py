"""A few Philox-mode g2048_step launches on 2^25 boards (for `ncu -k regex:step_lut_kernel`)."""
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
from bench import synthetic_packed  # noqa: E402
from simulate_2048_b200 import ops  # noqa: E402

ops.init()
n = 1 << 25
b = torch.from_numpy(np.tile(synthetic_packed(1 << 20), n >> 20).view(np.int64)).cuda()
a = torch.randint(0, 4, (n,), dtype=torch.uint8, device='cuda')
out = ops.step(b, a, None, seed=1)
for i in range(4):
    ops.step(b, a, None, seed=1, step_base=i, out=out)
torch.cuda.synchronize()
print('ok')
