"""A few from-reset rollouts of 2^22 envs (for `ncu -k regex:rollout_kernel`)."""
import sys

import torch

sys.path.insert(0, '.')
from simulate_2048_b200 import ops  # noqa: E402

ops.init()
n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 22)
for _ in range(3):
    r = ops.rollout(n, seed=1, per_env=False)
torch.cuda.synchronize()
print('steps', int(r.stats[1].item()))
