"""Host constants of render.py's contracted-scene sampling (render.py:127-155).  NumPy only (no torch): shared by the
torch front end (`tensorf_b200.render`) and the JAX binding (`jax_ffi/tensorf_jax.py`)."""
from __future__ import annotations

from typing import Tuple

import numpy as np


def contracted_schedule(near: float, far: float, n: int) -> Tuple[np.ndarray, np.ndarray]:
    """Host constants of the contracted branch (render.py:127-155): the `ts` schedule and step
    sizes, computed as the reference does (close_ts fp32 linspace, far_ts float64 numpy)."""
    nc = n // 2
    nf = n - nc
    f32 = np.float32
    if nc > 1:  # jnp.linspace [upstream]: start*(1-s_i) + stop*s_i, s_i = i/(num-1); endpoint appended exactly
        sv = np.arange(nc - 1, dtype=np.float32) * f32(1.0 / (nc - 1))
        close = (f32(near) * (f32(1.0) - sv) + f32(near + 1.0) * sv).astype(np.float32)
        close = np.concatenate([close, np.array([near + 1.0], dtype=np.float32)])
    else:
        close = np.full((nc,), near, dtype=np.float32)
    far_start = near + 1.0 + 1.0 / nc
    k = 10.0
    far_deltas = (1.0 / (1.0 - np.linspace(0.0, 1.0 - 1 / ((far - far_start) / k + 1), nf)) - 1.0) * np.linspace(1.0, k, nf)
    base = np.concatenate([close, (far_start + far_deltas).astype(np.float32)]).astype(np.float32)
    delta = np.roll(base, -1) - base
    delta[-1] = delta[-2]
    return base, delta.astype(np.float32)
