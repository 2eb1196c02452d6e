"""GPU parity of the n-step TD(lambda) return kernel."""

import numpy as np
import pytest

from oracle import np_engine as npe

pytestmark = pytest.mark.gpu

# f64 arithmetic on the device, f32 output: one rounding to float32 plus the difference between a
# running discount product and the reference's pow() — a few ulp of float32
RTOL = 4e-7


@pytest.fixture(scope='module')
def ops():
    from simulate_2048_b200 import ops
    ops.init()
    return ops


def test_nstep_returns_reproduce_reference(ops, golden):
    import torch
    g = golden('nstep')
    for i, (disc, n_steps, lam) in enumerate(g['params']):
        r = torch.from_numpy(g[f'r{i}']).cuda().reshape(-1, 1).contiguous()
        v = torch.from_numpy(g[f'v{i}']).cuda().reshape(-1, 1).contiguous()
        got = ops.nstep_returns(r, v, disc, int(n_steps), lam).cpu().numpy().reshape(-1)
        np.testing.assert_allclose(got, g[f'out{i}'], rtol=RTOL, atol=1e-6)


@pytest.mark.parametrize('T,N', [(1, 1), (17, 5), (64, 300), (200, 4096 + 3)])
def test_nstep_returns_batched_ragged_vs_oracle(ops, T, N):
    """Time-major [T,N] buffers with a different length per env; steps past the end are 0."""
    import torch
    rng = np.random.default_rng(T * 1000 + N)
    r = (rng.integers(0, 128, (T, N)) * 4).astype(np.float32)
    v = rng.normal(500.0, 200.0, (T, N)).astype(np.float32)
    lengths = rng.integers(0, T + 1, N).astype(np.int32)
    lengths[0] = T
    lengths[-1] = min(1, T)
    got = ops.nstep_returns(torch.from_numpy(r).cuda(), torch.from_numpy(v).cuda(), 0.997, 10, 0.5,
                            lengths=torch.from_numpy(lengths).cuda()).cpu().numpy()
    for e in rng.choice(N, size=min(N, 24), replace=False).tolist() + [0, N - 1]:
        L = int(lengths[e])
        want = npe.n_step_returns(r[:L, e].astype(np.float64), v[:L, e].astype(np.float64), 0.997, 10, 0.5)
        np.testing.assert_allclose(got[:L, e], want, rtol=RTOL, atol=1e-5)
        assert not got[L:, e].any()


def test_nstep_returns_argument_checks(ops):
    import torch
    r = torch.zeros((4, 3), dtype=torch.float32, device='cuda')
    with pytest.raises(ValueError):
        ops.nstep_returns(r, r[:, :2], 0.99, 5)
    with pytest.raises(ValueError):
        ops.nstep_returns(r.double(), r.double(), 0.99, 5)
    assert ops.nstep_returns(r, r, 0.99, 0).abs().sum().item() == 0.0  # no step: weights sum to 0 -> v[t]
