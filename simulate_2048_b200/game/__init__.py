"""Drop-in for ``reinforce.game`` (the JAX engine of simulate_2048) on B200.

Same names, argument order and return order as ``reinforce/game/core.py`` and
``reinforce/game/env.py``; ``jax.Array`` becomes a CUDA ``torch.Tensor`` and every call runs the
hand-written sm_100a kernels behind the C ABI.
"""

from .core import (  # noqa: F401
    count_empty,
    create_initial_board,
    encode_observation,
    fill_cells,
    is_done,
    latent_state,
    legal_actions_mask,
    max_tile,
    merge_row,
    next_state,
    pack_boards,
    slide_and_merge,
    unpack_boards,
)
from .env import (  # noqa: F401
    Environment,
    GameState,
    batched_get_legal_actions,
    batched_get_observation,
    batched_reset,
    batched_step,
    get_legal_actions,
    get_observation,
    reset,
    step,
)
