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


def test_sampled_children_vs_oracle(oracle):
    """g2048_sample_children: bit-exact against the oracle's draws (word of (seed, board, draw, tag 5),
    digits 1 and 2), every child is a member of the exact child set, illegal actions give zeros."""
    import torch
    from simulate_2048_b200 import ops
    from tests.helpers import synthetic_packed
    n, k = 3000, 5
    boards = synthetic_packed(n, seed=41, max_exp=8)
    d = torch.from_numpy(boards.view(np.int64)).cuda()
    ex = ops.expand(d, children='compact')
    kids = ops.sample_children(ex, k, seed=99, env_offset=7, draw_base=40)
    want = oracle.sample_children(ex.after.cpu().numpy().view(np.uint64), ex.child_count.cpu().numpy(), k,
                                  seed=99, env_offset=7, draw_base=40)
    got = kids.cpu().numpy().view(np.uint64)
    np.testing.assert_array_equal(got, want)
    cc, offs = ex.child_count.cpu().numpy(), ex.child_offsets.cpu().numpy()
    allkids = ex.children.cpu().numpy().view(np.uint64)
    for i in range(0, n, 37):
        o = int(offs[i])
        for a in range(4):
            c = int(cc[i, a])
            if c == 0:
                assert (got[i, a] == 0).all()
            else:
                assert set(got[i, a].tolist()) <= set(allkids[o:o + c].tolist())
            o += c


def test_progressive_widening_converges_to_exact():
    """expectimax(widen=k): an unbiased Monte-Carlo estimate of the exact backup — the error against
    widen=None shrinks like 1/sqrt(k), the greedy action mostly agrees at k = 64, and depth 2 with
    k = 4 expands 16x fewer boards per level than the exact tree."""
    import torch
    from simulate_2048_b200 import ops, selfplay
    from tests.helpers import synthetic_packed
    boards = torch.from_numpy(synthetic_packed(4000, seed=3, max_exp=7).view(np.int64)).cuda()

    def value(children):
        obs, mx = ops.observe(children)
        return (obs == 0).sum(dim=1).to(torch.float32) * 3.0 + mx.to(torch.float32)
    exact, best, _ = selfplay.expectimax(boards, value, depth=1)
    legal = ~torch.isinf(exact)
    errs = []
    for k in (4, 16, 64):
        q, _, _ = selfplay.expectimax(boards, value, depth=1, widen=k, seed=5)
        assert torch.equal(torch.isinf(q), torch.isinf(exact))
        errs.append(float((q[legal] - exact[legal]).abs().mean()))
    assert errs[0] > errs[1] > errs[2] and errs[2] < 0.45 * errs[0]
    q64, best64, _ = selfplay.expectimax(boards, value, depth=1, widen=64, seed=5)
    has = legal.any(dim=1)
    assert float((best64[has] == best[has]).float().mean()) > 0.9
    bias = float((q64[legal] - exact[legal]).mean())
    assert abs(bias) < 0.05                                   # unbiased
    q2, b2, _ = selfplay.expectimax(boards[:512].contiguous(), value, depth=2, widen=4, seed=1)
    e2, be2, _ = selfplay.expectimax(boards[:512].contiguous(), value, depth=2)
    ok = ~torch.isinf(e2)
    assert torch.equal(torch.isinf(q2), torch.isinf(e2))
    assert float((q2[ok] - e2[ok]).abs().mean()) < 0.25 * float(e2[ok].abs().mean())
