"""Device time and algorithmic GB/s of the supporting kernels at 2^24 boards (CUDA events around five
back-to-back calls, median of 7 samples)."""
import statistics
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
from bench import synthetic_packed  # noqa: E402
from simulate_2048_b200 import ops  # noqa: E402

ops.init()
n = 1 << 24
b = torch.from_numpy(np.tile(synthetic_packed(1 << 20), n >> 20).view(np.int64)).cuda()
acts = torch.randint(0, 4, (n,), dtype=torch.uint8, device='cuda')


def timed(name, fn, bytes_per_board):
    fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(7):   # 6 back-to-back calls per sample: the queue stays full, host overhead drops out
        e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        fn()
        e[0].record()
        for _ in range(5):
            fn()
        e[1].record()
        torch.cuda.synchronize()
        ms.append(e[0].elapsed_time(e[1]) / 5)
    t = statistics.median(ms)
    print(f'{name:28s} {t * 1e3:9.1f} us  {n * bytes_per_board / t / 1e6:8.0f} GB/s algorithmic (incl. allocation of the outputs)')


timed('legal_flags (mask)', lambda: ops.legal_flags(b), 9)
timed('observe (+max)', lambda: ops.observe(b), 8 + 64 + 1)
timed('observe_onehot w=31', lambda: ops.observe_onehot(b[:n // 4]), (8 + 496) / 4)
tiles = ops.unpack(b)
timed('unpack', lambda: ops.unpack(b), 72)
timed('pack', lambda: ops.pack(tiles), 72)
timed('reset (Philox)', lambda: ops.reset(n, seed=3), 8)
timed('spawn (Philox)', lambda: ops.spawn(b.clone(), None, seed=5), 16 + 16)
timed('after_state (dense 32)', lambda: ops.after_state(b[:n // 8]), (8 + 256 + 3) / 8)
st = torch.zeros(n, dtype=torch.int32, device='cuda')
dn = torch.zeros(n, dtype=torch.uint8, device='cuda')
tr = torch.zeros(n, dtype=torch.float32, device='cuda')
bb = b.clone()
timed('env_step (state only)', lambda: ops.env_step(bb, st, dn, tr, acts, None, seed=1), 2 * (8 + 4 + 1 + 4) + 1)
# round 2 kernels
timed('legal_actions (bool[N,4])', lambda: ops.legal_actions(b), 8 + 4)
pol = torch.rand((n, 4), device='cuda')
legal = ops.legal_actions(b)
timed('sample_action', lambda: ops.sample_action(pol, legal, 1.0, seed=3), 16 + 4 + 1)
obs = torch.empty((n, 16), dtype=torch.float32, device='cuda')
lg = torch.empty((n, 4), dtype=torch.bool, device='cuda')
rw = torch.empty(n, dtype=torch.float32, device='cuda')
timed('selfplay_step (no record)', lambda: ops.selfplay_step(bb, st, dn, tr, pol, temperature=1.0, seed=1, rewards=rw, obs=obs, legal=lg),
      2 * (8 + 4 + 1 + 4) + 16 + 4 + 64 + 4)
m = n // 16
ex = ops.expand(b[:m].contiguous(), children='none')
timed('sample_children k=4', lambda: ops.sample_children(ex, 4, seed=1), (32 + 4 + 16 * 8) / 16)
vals = torch.rand(m * 16, device='cuda')
timed('expect_sampled k=4', lambda: ops.expect_sampled(vals, ex, 4), (64 + 4 + 16 + 16) / 16)
