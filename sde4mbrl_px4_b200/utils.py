"""``enu2ned`` / ``ned2enu`` and the opaque-key helpers of the controller interface.

The reference imports ``enu2ned`` from the un-vendored package
(``from sde4mbrlExamples.rotor_uav.utils import enu2ned``, sde_control.py:13) and
calls it as ``enu2ned(curr_state, np)`` (sde_control.py:400).  Semantics per
SURVEY.md section 8a [SPEC] "State / frames".
"""
from __future__ import annotations

import numpy as _np

_S = _np.float32(0.70710678118654752440)


def enu2ned(x, xp=_np):
    """ENU/FLU 13-state -> NED/FRD 13-state (also its own inverse).

    p, v: (x,y,z)->(y,x,-z); w: (x,y,z)->(x,-y,-z);
    q_ned = q_r (x) q_enu (x) q_b with q_r=(0,s,s,0), q_b=(0,1,0,0), sign-fixed to qw>=0.
    ``xp`` is the array namespace (the reference passes ``np``).  Works on [..., 13].
    """
    x = xp.asarray(x)
    qw, qx, qy, qz = x[..., 6], x[..., 7], x[..., 8], x[..., 9]
    c0, c1, c2, c3 = _S * (qw + qz), _S * (qx + qy), _S * (qx - qy), _S * (qw - qz)
    sg = xp.where(c0 < 0, -1.0, 1.0).astype(x.dtype)
    cols = [x[..., 1], x[..., 0], -x[..., 2], x[..., 4], x[..., 3], -x[..., 5],
            sg * c0, sg * c1, sg * c2, sg * c3, x[..., 10], -x[..., 11], -x[..., 12]]
    return xp.stack(cols, axis=-1).astype(x.dtype)


ned2enu = enu2ned


def PRNGKey(seed: int) -> _np.ndarray:
    """Opaque solver key: uint64[2] = (seed, tick counter).  Mirrors
    ``jax.random.PRNGKey(self.seed)`` (sde_control.py:338); the generator behind it is
    Philox4x32-10, not threefry, so streams differ from JAX by construction."""
    return _np.array([_np.uint64(seed), _np.uint64(0)], dtype=_np.uint64)


def split(key, num: int = 2):
    """Derive ``num`` independent keys (mirrors ``jax.random.split``, sde_control.py:341):
    seed_i = splitmix64(seed + i + 1), tick counter reset to 0."""
    key = _np.asarray(key, dtype=_np.uint64)
    out = _np.zeros((num, 2), _np.uint64)
    m = (1 << 64) - 1
    for i in range(num):
        z = (int(key[0]) + (i + 1) * 0x9E3779B97F4A7C15) & m
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & m
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & m
        out[i, 0] = _np.uint64(z ^ (z >> 31))
    return out
