"""Times g2048_rollout (config 4) at several batch sizes: event-timed, second call of each size."""
import sys

import torch

sys.path.insert(0, '.')
from simulate_2048_b200 import ops  # noqa: E402

ops.init()
for lg in (20, 22, 24):
    n = 1 << lg
    for _ in range(2):
        ops.rollout(n, seed=1, per_env=False)
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    e[0].record()
    r = ops.rollout(n, seed=1, per_env=False)
    e[1].record()
    torch.cuda.synchronize()
    steps = int(r.stats[1].item())
    ms = e[0].elapsed_time(e[1])
    print(n, round(ms, 2), 'ms', round(steps / ms / 1e6, 1), 'G env-steps/s')
