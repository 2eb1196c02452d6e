"""Depth-1 expectimax backup (g2048_expect) against values computed with the
real reference's latent_state / after_state probabilities (tests/golden/expectimax.npz) and against
the oracle restatement on other boards."""

import numpy as np
import pytest

from oracle import np_engine as npe
from tests.helpers import synthetic_packed

pytestmark = pytest.mark.gpu

# f32 accumulation of at most 30 products on the device against the reference's f64 sum
RTOL, ATOL = 2e-6, 1e-4


def leaf_value_np(tiles) -> float:
    t = np.asarray(tiles)
    mx = int(t.max())
    return float((t == 0).sum()) + 0.25 * (mx.bit_length() - 1 if mx else 0)


def leaf_value_gpu(children):
    """Same leaf value on packed boards: empty cells + max exponent / 4."""
    import torch
    from simulate_2048_b200 import ops
    obs, mx = ops.observe(children)
    return (obs == 0).sum(dim=1).to(torch.float32) + 0.25 * mx.to(torch.float32)


def test_expectimax_reproduces_reference(golden):
    import torch
    from simulate_2048_b200.selfplay import expectimax_depth1
    g = golden('expectimax')
    boards = torch.from_numpy(g['packed'].view(np.int64)).cuda()
    q, best, ex = expectimax_depth1(boards, leaf_value_gpu, discount=float(g['discount']))
    got, want = q.cpu().numpy().astype(np.float64), g['q']
    assert np.array_equal(np.isinf(got), np.isinf(want))
    fin = np.isfinite(want)
    np.testing.assert_allclose(got[fin], want[fin], rtol=RTOL, atol=ATOL)
    has_move = fin.any(axis=1)
    assert np.array_equal(best.cpu().numpy()[has_move], np.argmax(want, axis=1)[has_move]) or \
        np.allclose(np.take_along_axis(want, best.cpu().numpy()[:, None], 1)[has_move, 0], want.max(axis=1)[has_move], rtol=1e-5)


@pytest.mark.parametrize('n', [1, 33, 5000])
def test_expectimax_vs_oracle(oracle, n):
    import torch
    from simulate_2048_b200.selfplay import expectimax_depth1
    packed = synthetic_packed(n, seed=n + 9, max_exp=9)
    tiles = oracle.unpack(packed).astype(np.int64).reshape(n, 4, 4)
    q, _, _ = expectimax_depth1(torch.from_numpy(packed.view(np.int64)).cuda(), leaf_value_gpu, discount=0.9)
    got = q.cpu().numpy().astype(np.float64)
    for i in np.random.default_rng(n).choice(n, size=min(n, 64), replace=False):
        want = npe.expectimax_q(tiles[i], leaf_value_np, discount=0.9)
        assert np.array_equal(np.isinf(got[i]), np.isinf(want))
        fin = np.isfinite(want)
        np.testing.assert_allclose(got[i][fin], want[fin], rtol=RTOL, atol=ATOL)


def test_expect_argument_checks():
    import torch
    from simulate_2048_b200 import ops
    boards = torch.from_numpy(synthetic_packed(8, seed=1).view(np.int64)).cuda()
    dense = ops.expand(boards, children='dense')
    with pytest.raises(ValueError):
        ops.expect(torch.zeros(10, device='cuda'), dense)
    compact = ops.expand(boards, children='compact')
    with pytest.raises(ValueError):
        ops.expect(torch.zeros(3, device='cuda'), compact)


def test_expectimax_depth2_reproduces_reference(golden):
    """Two levels: V(child) = max over its legal actions of the depth-1 value (0 for a dead child),
    computed by the fixture with the real reference's functions."""
    import torch
    from simulate_2048_b200.selfplay import expectimax
    g = golden('expectimax')
    boards = torch.from_numpy(g['deep_packed'].view(np.int64)).cuda()
    q, _, _ = expectimax(boards, leaf_value_gpu, depth=2, discount=float(g['discount']))
    got, want = q.cpu().numpy().astype(np.float64), g['deep_q']
    assert np.array_equal(np.isinf(got), np.isinf(want))
    fin = np.isfinite(want)
    np.testing.assert_allclose(got[fin], want[fin], rtol=5e-6, atol=1e-3)
