"""Seeded synthetic problem generators for the BASELINE configs."""
from __future__ import annotations

import numpy as np

from . import trajectory


def initial_states(p_ref0: np.ndarray, n: int, seed: int = 0) -> np.ndarray:
    """p = p_ref(0)+U(-0.5,0.5)^3, v~N(0,0.3^2), q = normalise((1,0,0,0)+N(0,0.05^2)),
    w~N(0,0.1^2); ENU, float32 [n,13]."""
    rng = np.random.default_rng(seed)
    x = np.zeros((n, 13), np.float64)
    x[:, 0:3] = np.asarray(p_ref0, np.float64).reshape(-1, 3) + rng.uniform(-0.5, 0.5, (n, 3))
    x[:, 3:6] = rng.normal(0.0, 0.3, (n, 3))
    q = np.array([1.0, 0, 0, 0]) + rng.normal(0.0, 0.05, (n, 4))
    x[:, 6:10] = q / np.linalg.norm(q, axis=1, keepdims=True)
    x[:, 10:13] = rng.normal(0.0, 0.1, (n, 3))
    return x.astype(np.float32)


def batched_problems(B: int, horizon: int, dt: np.ndarray, seed: int = 0):
    """B independent iris problems: per-problem random lemniscate reference window
    (explicit xref_win, ENU), random initial state around the window start, rng seed 1000+b.
    Returns dict(x[B,13], xref_win[B,H+1,13], rng[B,2])."""
    rng = np.random.default_rng(seed)
    tgrid = np.concatenate([[0.0], np.cumsum(np.asarray(dt[:horizon], np.float64))])
    win = np.zeros((B, horizon + 1, 13), np.float32)
    for b in range(B):
        A, Tp, ph = rng.uniform(1.0, 3.0), rng.uniform(6.0, 12.0), rng.uniform(0.0, 2.0 * np.pi)
        t0 = rng.uniform(0.0, Tp)
        w = 2.0 * np.pi / Tp
        a = w * (t0 + tgrid) + ph
        s, c = np.sin(a), np.cos(a)
        win[b, :, 0], win[b, :, 1], win[b, :, 2] = A * s, A * s * c, 1.5
        win[b, :, 3], win[b, :, 4] = A * w * c, A * w * (c * c - s * s)
        win[b, :, 6] = 1.0
    x = initial_states(win[:, 0, 0:3], B, seed + 1)
    r = np.zeros((B, 2), np.uint64)
    r[:, 0] = 1000 + np.arange(B)
    return dict(x=x, xref_win=win, rng=r)
