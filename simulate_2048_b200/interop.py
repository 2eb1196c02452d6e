"""Zero-copy hand-off of device buffers to other frameworks.

The reference's learner is JAX (``reinforce/training``); its arrays and ours can share device memory
through DLPack (``jax.dlpack.from_dlpack`` on the consumer side), so observations and legal masks
produced by the kernels reach the network without a copy, and the network's policy output can be
written straight into ``BatchedEnv.policy_in``.  Nothing here copies or computes.
"""

from __future__ import annotations

from typing import Dict

import torch
from torch.utils import dlpack as _dl

from .game.env import GameState


def to_dlpack(t: torch.Tensor):
    """DLPack capsule of a CUDA tensor (the producer keeps ownership until the consumer drops it)."""
    if not t.is_cuda:
        raise ValueError('to_dlpack: CUDA tensor required (simulate_2048_b200 has no CPU path)')
    return _dl.to_dlpack(t)


def from_dlpack(capsule_or_array) -> torch.Tensor:
    """CUDA tensor that aliases a DLPack capsule or any object with ``__dlpack__`` (e.g. a jax.Array)."""
    return _dl.from_dlpack(capsule_or_array)


def state_to_dlpack(state: GameState) -> Dict[str, object]:
    """One capsule per GameState field (env.py:33): board (packed int64), step_count, done (uint8
    view of the bool tensor: DLPack has no portable bool before v1.0), total_reward."""
    return {'board': to_dlpack(state.board), 'step_count': to_dlpack(state.step_count),
            'done': to_dlpack(state.done.view(torch.uint8)), 'total_reward': to_dlpack(state.total_reward)}


def state_from_dlpack(capsules: Dict[str, object]) -> GameState:
    return GameState(from_dlpack(capsules['board']), from_dlpack(capsules['step_count']),
                     from_dlpack(capsules['done']).view(torch.bool), from_dlpack(capsules['total_reward']))
