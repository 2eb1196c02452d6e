"""Host-side argument handling of the API mirrors (no GPU needed)."""

import pytest
import torch

from simulate_2048_b200.game import core, env


def test_key_to_seed_forms():
    assert core.key_to_seed(None) == 0 and core.key_to_seed(42) == 42
    assert core.key_to_seed(-1) == 2**64 - 1
    assert core.key_to_seed(torch.tensor(7)) == 7
    assert core.key_to_seed(torch.tensor([1, 2], dtype=torch.int64)) == 1 | (2 << 32)
    with pytest.raises(ValueError):
        core.key_to_seed(torch.zeros(3))


def test_split_keys_dispatch():
    n = 5
    assert env._split_keys(None, n) == (0, None, None)
    assert env._split_keys(9, n) == (9, None, None)
    seed, keys, spawns = env._split_keys(torch.zeros(n, dtype=torch.uint8), n)
    assert keys is None and spawns.numel() == n
    seed, keys, spawns = env._split_keys(torch.arange(n, dtype=torch.int64), n)
    assert spawns is None and keys.tolist() == list(range(n))
    jax_keys = torch.tensor([[1, 2]] * n, dtype=torch.int64)
    seed, keys, spawns = env._split_keys(jax_keys, n)
    assert keys.tolist() == [1 | (2 << 32)] * n


def test_is_packed_detection():
    assert core.is_packed(torch.zeros(7, dtype=torch.int64))
    assert not core.is_packed(torch.zeros((7, 4, 4), dtype=torch.int32))
    assert not core.is_packed(torch.zeros((4, 4), dtype=torch.int32))
    with pytest.raises(TypeError):
        core.pack_boards(torch.zeros(3, dtype=torch.int32))


def test_bench_rank_local_block_has_no_collective():
    """bench.py's secondary measurements run on rank 0 only: a barrier / all-reduce inside that
    block pairs with the other ranks' final barrier and hangs the torchrun job (seen at N=2)."""
    import os
    import re
    src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'bench.py')).read()
    a = src.index('# ---- secondary measurements (rank-local, short) ----')
    b = src.index("line['secondary'] = extra")
    block = src[a:b]
    assert 'barrier(' not in block and 'max_over_ranks(' not in block and 'dist.' not in block
    calls = re.findall(r'timed\(.*', block)
    assert calls and all('collective=False' in c for c in calls), calls
