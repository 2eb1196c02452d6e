"""Small driver for profiling the rollout (config 4) and expand (config 3) kernels under ncu."""
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
from bench import synthetic_packed  # noqa: E402
from simulate_2048_b200 import ops  # noqa: E402

ops.init()
boards = torch.from_numpy(synthetic_packed(1 << 20).view(np.int64)).cuda()
ex = ops.expand(boards, children='compact')
for _ in range(2):
    ops.expand(boards, children='compact', out=ex)
    ops.rollout(1 << 22, seed=1, per_env=False)
torch.cuda.synchronize()
print('children', ex.children.numel())
