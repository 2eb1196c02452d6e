"""Shared test helpers: the synthetic board generator of SURVEY.md §8d and packing."""

import numpy as np


def synthetic_exponents(n, seed=2048, max_exp=11):
    """n_tiles ~ U{1..16}, positions = first n of a random order, exponents ~ U{1..max_exp};
    draws in the order integers, random, integers (tests/test_perf_utils.py:12-19 vectorised)."""
    rng = np.random.default_rng(seed)
    cnt = rng.integers(1, 17, n)
    keys = rng.random((n, 16))
    ex = rng.integers(1, max_exp + 1, (n, 16))
    order = np.argsort(keys, axis=1, kind='stable')
    ranks = np.empty_like(order)
    np.put_along_axis(ranks, order, np.arange(16)[None, :].repeat(n, 0), axis=1)
    # cell c receives ex[i, ranks[i, c]] when ranks[i, c] < cnt[i]
    vals = np.take_along_axis(ex, ranks, axis=1)
    return np.where(ranks < cnt[:, None], vals, 0).astype(np.uint64)


def pack_exponents(exps):
    shifts = (4 * np.arange(16, dtype=np.uint64))[None, :]
    return np.bitwise_or.reduce(exps.astype(np.uint64) << shifts, axis=1)


def synthetic_packed(n, seed=2048, max_exp=11):
    return pack_exponents(synthetic_exponents(n, seed, max_exp))


def pack_tiles(tiles):
    """int tile values [..., 4, 4] -> packed u64[...] (host helper for tests only)."""
    t = np.asarray(tiles, dtype=np.int64).reshape(-1, 16)
    e = np.zeros_like(t)
    nz = t > 0
    e[nz] = np.round(np.log2(t[nz])).astype(np.int64)
    return pack_exponents(e.astype(np.uint64)).reshape(np.asarray(tiles).shape[:-2])


def random_spawn_bytes(n, n_empty_hint=None, seed=3):
    """rank ~ U{0..15} (tests clamp by the true empties via modulo), four w.p. 0.1"""
    rng = np.random.default_rng(seed)
    rank = rng.integers(0, 16, n).astype(np.uint8)
    four = (rng.random(n) < 0.1).astype(np.uint8)
    return rank, four


def valid_spawn_bytes(after_boards, seed=3):
    """Spawn bytes whose rank is uniform over the true empties of the given afterstates."""
    b = np.asarray(after_boards, dtype=np.uint64)
    shifts = (4 * np.arange(16, dtype=np.uint64))[None, :]
    empt = (((b[:, None] >> shifts) & np.uint64(0xF)) == 0).sum(axis=1)
    rng = np.random.default_rng(seed)
    rank = (rng.integers(0, 1 << 30, b.size) % np.maximum(empt, 1)).astype(np.uint8)
    four = (rng.random(b.size) < 0.1).astype(np.uint8)
    return (rank | (four << 4)).astype(np.uint8)
