"""Host-side logic of the multi-GPU path on CPU: env sharding and the statistics collective,
world_size 2 over gloo.  The oracle stands in for the kernels (tests may use it), which also
shows that sharded rollouts are bit-identical to the single-rank run because the Philox stream
is keyed by the GLOBAL env id."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from simulate_2048_b200.parallel import EpisodeStats, combine_stats, reduce_stats, shard_range


def test_shard_range_partitions():
    for n in (0, 1, 7, 1 << 20, (1 << 24) + 5):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def test_combine_stats_sum_and_max():
    a = torch.arange(24, dtype=torch.int64)
    b = torch.arange(24, dtype=torch.int64) * 2
    c = combine_stats(torch.stack([a, b]))
    assert c[0] == 0 and c[1] == 3 and c[3] == 6 and c[4] == 8 and c[5] == 15 and c[23] == 69


def test_episode_stats_summary():
    v = torch.zeros(24, dtype=torch.int64)
    v[0], v[1], v[2], v[3], v[4] = 10, 1170, 11000, 3000, 250
    v[8 + 7], v[8 + 11] = 9, 1
    s = EpisodeStats.from_vector(v).summary()
    assert s['mean_length'] == 117 and s['mean_reward'] == 1100 and s['max_tile'] == 2048
    assert s['reached_2048'] == 1 and s['reached_4096'] == 0 and s['mean_max_tile'] == (9 * 128 + 2048) / 10


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, seed, out_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from oracle import c_oracle
    lo, hi = shard_range(n, rank, world)
    local = c_oracle.rollout(hi - lo, seed=seed, env_offset=lo)
    stats = torch.from_numpy(local['stats'].view(np.int64).copy())
    total = reduce_stats(stats)
    np.save(os.path.join(out_dir, f'stats{rank}.npy'), total.numpy())
    np.save(os.path.join(out_dir, f'score{rank}.npy'), local['score'])
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_rollout_world2_gloo(tmp_path):
    from oracle import c_oracle
    c_oracle.build()
    n, seed, world = 301, 17, 2
    mp.spawn(_worker, args=(world, _free_port(), n, seed, str(tmp_path)), nprocs=world, join=True)
    whole = c_oracle.rollout(n, seed=seed)
    s0, s1 = np.load(tmp_path / 'stats0.npy'), np.load(tmp_path / 'stats1.npy')
    np.testing.assert_array_equal(s0, s1)  # every rank holds the global vector
    np.testing.assert_array_equal(s0.view(np.uint64), whole['stats'])
    scores = np.concatenate([np.load(tmp_path / 'score0.npy'), np.load(tmp_path / 'score1.npy')])
    np.testing.assert_array_equal(scores, whole['score'])


def test_reduce_stats_without_process_group():
    v = torch.arange(24, dtype=torch.int64)
    assert torch.equal(reduce_stats(v), v)
