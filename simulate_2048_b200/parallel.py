"""Env sharding over the GPUs of one node (one process per GPU, ``torch.distributed``).

Environments are independent, so the step path has NO collective: rank ``r`` of ``W`` owns the
contiguous env range ``shard_range(n, r, W)`` and passes its start as ``env_offset``, which keys
the Philox stream by the GLOBAL env id — results are identical for every ``W``.  The only
exchange is one small collective per rollout for the episode statistics (the
reference reports the same quantities in ``self_play.evaluate_games``, self_play.py:503-521).

The host logic here is device-agnostic (it only touches the 24-word statistics vector), so it is
covered on CPU with the gloo backend; the kernels themselves only run on CUDA.
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import torch
import torch.distributed as dist

STATS_WORDS = 24
_MAX_WORDS = (3, 4)  # max score, max steps: combined with max, everything else with sum


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced split: sizes differ by at most one env."""
    if not 0 <= rank < world:
        raise ValueError(f'rank {rank} outside world of {world}')
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def combine_stats(stack: torch.Tensor) -> torch.Tensor:
    """[W, 24] per-rank statistics -> [24] global (sum, except the two max words)."""
    out = stack.sum(dim=0)
    lo, hi = _MAX_WORDS[0], _MAX_WORDS[-1] + 1          # the two max words are adjacent
    out[lo:hi] = stack[:, lo:hi].amax(dim=0)
    return out


def reduce_stats(stats: torch.Tensor, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """One collective (all-gather of 24 int64 words per rank), combined locally.  Returns the
    global vector on every rank; a no-op when torch.distributed is not initialised."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return stats.clone()
    world = dist.get_world_size(group)
    flat = torch.empty(world * STATS_WORDS, dtype=stats.dtype, device=stats.device)
    dist.all_gather_into_tensor(flat, stats.contiguous().reshape(-1), group=group)
    return combine_stats(flat.reshape(world, STATS_WORDS))


@dataclass
class EpisodeStats:
    """Decoded statistics vector (include/g2048.h g2048_rollout)."""

    episodes: int
    total_steps: int
    total_score: int
    max_score: int
    max_steps: int
    capped: int
    overflow: int
    max_tile_hist: Dict[int, int]
    vector: Tuple[int, ...] = ()   # the raw 24 words

    @classmethod
    def from_vector(cls, v) -> 'EpisodeStats':
        w = [int(x) for x in v.tolist()]
        hist = {(1 << e if e else 0): w[8 + e] for e in range(16) if w[8 + e]}
        return cls(w[0], w[1], w[2], w[3], w[4], w[5], w[6], hist, tuple(w))

    def summary(self) -> Dict[str, float]:
        """The quantities self_play.evaluate_games reports (self_play.py:503-521)."""
        n = max(self.episodes, 1)
        out = {
            'episodes': self.episodes,
            'mean_reward': self.total_score / n,
            'max_reward': float(self.max_score),
            'mean_length': self.total_steps / n,
            'max_length': self.max_steps,
            'max_tile': max(self.max_tile_hist) if self.max_tile_hist else 0,
            'mean_max_tile': sum(t * c for t, c in self.max_tile_hist.items()) / n,
        }
        for tile in (2048, 4096, 8192, 16384, 32768):
            out[f'reached_{tile}'] = sum(c for t, c in self.max_tile_hist.items() if t >= tile)
        return out


def sharded_rollout(n_total: int, *, seed: int = 0, max_steps: int = 1 << 20,
                    rank: Optional[int] = None, world: Optional[int] = None,
                    per_env: bool = False, group: Optional[dist.ProcessGroup] = None, device=None,
                    to_host: bool = True):
    """Random rollouts of envs [0, n_total) split over the ranks.  Returns (local RolloutOut,
    global EpisodeStats); with ``to_host=False`` the second value is the reduced 24-word vector as
    a DEVICE tensor (no host synchronisation: decode it later with ``EpisodeStats.from_vector``)."""
    from . import ops  # CUDA only

    if rank is None or world is None:
        init = dist.is_available() and dist.is_initialized()
        rank = dist.get_rank(group) if init else 0
        world = dist.get_world_size(group) if init else 1
    lo, hi = shard_range(n_total, rank, world)
    out = ops.rollout(hi - lo, seed=seed, env_offset=lo, max_steps=max_steps, per_env=per_env,
                      device=device)
    reduced = reduce_stats(out.stats, group)
    return out, (EpisodeStats.from_vector(reduced.cpu()) if to_host else reduced)
