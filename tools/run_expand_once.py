"""A few compact expansions of 65536 boards into preallocated buffers (for ncu)."""
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
from bench import synthetic_packed  # noqa: E402
from simulate_2048_b200 import ops  # noqa: E402

ops.init()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
b = torch.from_numpy(synthetic_packed(1 << 20)[:n].view(np.int64).copy()).cuda()
ex = ops.expand(b, children='compact', capacity=72 * n)
for _ in range(4):
    ops.expand(b, children='compact', out=ex)
dn = ops.expand(b, children='dense')
for _ in range(2):
    ops.expand(b, children='dense', out=dn)
torch.cuda.synchronize()
print('total', ex.total.tolist())
