"""Times g2048_rollout (config 4) at several batch sizes: CUDA-event timed, min and median of 5 calls."""
import statistics
import sys

import torch

sys.path.insert(0, '.')
from simulate_2048_b200 import ops  # noqa: E402

ops.init()
for lg in (20, 21, 22, 24):
    n = 1 << lg
    ops.rollout(n, seed=1, per_env=False)
    torch.cuda.synchronize()
    ms = []
    for _ in range(5):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        e[0].record()
        r = ops.rollout(n, seed=1, per_env=False)
        e[1].record()
        torch.cuda.synchronize()
        ms.append(e[0].elapsed_time(e[1]))
    steps = int(r.stats[1].item())
    print(n, 'min', round(min(ms), 2), 'ms', round(steps / min(ms) / 1e6, 1), 'G env-steps/s; median',
          round(steps / statistics.median(ms) / 1e6, 1))
