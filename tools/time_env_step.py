"""Times BatchedEnv.step (g2048_env_step, one launch per iteration) at several batch sizes."""
import sys

import torch

sys.path.insert(0, '.')
from simulate_2048_b200 import ops  # noqa: E402
from simulate_2048_b200.selfplay import BatchedEnv  # noqa: E402

ops.init()
for log2 in (12, 16, 20, 22):
    n = 1 << log2
    env = BatchedEnv(n, seed=1, graph=True)
    env.reset()
    actions = torch.randint(0, 4, (n,), dtype=torch.uint8, device='cuda')
    for _ in range(5):
        env.step(actions)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(50):
        env.step(actions)
    ev[1].record()
    torch.cuda.synchronize()
    us = ev[0].elapsed_time(ev[1]) * 1000 / 50
    print(n, round(us, 1), 'us/iter', round(n / us / 1e3, 2), 'G env-steps/s', round(n * 106 / us / 1e3, 1), 'GB/s')
