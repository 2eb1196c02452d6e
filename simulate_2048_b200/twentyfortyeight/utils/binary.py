"""Drop-in for ``twentyfortyeight/utils/binary.py``: one-hot log2 encodings, computed by
``g2048_observe_onehot`` on the device."""

from __future__ import annotations

import numpy as np
from numpy import int64, ndarray

from ... import ops
from ..core._device import to_packed


def encode(state: ndarray, encodage_size: int) -> ndarray:
    """binary.py:11 — ``eye(encodage_size)[log2(state)]`` for a flat array of 16 tile values
    (0 encodes as index 0).  Shape (state.size, encodage_size), dtype int64."""
    flat = np.asarray(state).reshape(-1)
    if flat.size != 16:
        raise NotImplementedError('only the 16 cells of the 4x4 board are implemented on the device path')
    onehot = ops.observe_onehot(to_packed(flat.reshape(4, 4)), width=int(encodage_size))
    return onehot[0].cpu().numpy().astype(int64)


def encode_flatten(state: ndarray, encodage_size: int) -> ndarray:
    """binary.py:52 — flatten, then :func:`encode`, then flatten again (length 16 * encodage_size)."""
    return encode(np.asarray(state).ravel(), encodage_size).ravel()
