"""Reference trajectories (``trajectory_path`` of the YAML, launch/iris_sitl_traj_mpc.yaml:6).

CSV column contract ``t,x,y,z,vx,vy,vz,ax,ay,az,yaw`` is the one the reference's
geometric controller reads (geometric_controller.cpp:463); no CSV ships with the
reference, so a seeded lemniscate generator (name follows ``fast2_lemn.csv``) is
provided.
"""
from __future__ import annotations

import os

import numpy as np

CSV_COLUMNS = ("t", "x", "y", "z", "vx", "vy", "vz", "ax", "ay", "az", "yaw")


def lemniscate(amplitude: float = 2.0, period: float = 8.0, phase: float = 0.0, height: float = 1.5,
               duration: float = 30.0, rate_hz: float = 100.0) -> np.ndarray:
    """Lemniscate of Gerono in ENU sampled at ``rate_hz``: x = A sin(wt+ph),
    y = A sin(wt+ph) cos(wt+ph), z = height, yaw = 0.  Returns [T, 11] float64 in
    the CSV column order."""
    n = int(round(duration * rate_hz)) + 1
    t = np.arange(n) / rate_hz
    w = 2.0 * np.pi / period
    a = w * t + phase
    s, c = np.sin(a), np.cos(a)
    x, y = amplitude * s, amplitude * s * c
    vx, vy = amplitude * w * c, amplitude * w * (c * c - s * s)
    ax, ay = -amplitude * w * w * s, -4.0 * amplitude * w * w * s * c
    z = np.full_like(t, height)
    zero = np.zeros_like(t)
    return np.stack([t, x, y, z, vx, vy, zero, ax, ay, zero, zero], axis=1)


def save_csv(path: str, rows: np.ndarray) -> None:
    np.savetxt(path, rows, delimiter=",", header=",".join(CSV_COLUMNS), comments="", fmt="%.9g")


def load_csv(path: str) -> np.ndarray:
    """Read a trajectory CSV with a header row naming at least the CSV_COLUMNS
    (extra columns are ignored, order is free).  Returns [T, 11] float64."""
    path = os.path.expanduser(path)
    with open(path, "r") as f:
        header = [h.strip() for h in f.readline().strip().split(",")]
    missing = [c for c in CSV_COLUMNS if c not in header]
    if missing:
        raise ValueError(f"{path}: missing trajectory columns {missing}; expected {CSV_COLUMNS}")
    data = np.loadtxt(path, delimiter=",", skiprows=1, ndmin=2)
    return data[:, [header.index(c) for c in CSV_COLUMNS]]


def csv_rows_to_table(rows: np.ndarray) -> np.ndarray:
    """[T, 11] CSV rows -> [T, 14] float32 table (t + 13-state, ENU/FLU):
    quaternion from yaw only, reference body rates zero (SURVEY.md 8a [SPEC]
    "Reference window")."""
    rows = np.asarray(rows, np.float64)
    T = rows.shape[0]
    tab = np.zeros((T, 14), np.float64)
    tab[:, 0] = rows[:, 0]
    tab[:, 1:4] = rows[:, 1:4]
    tab[:, 4:7] = rows[:, 4:7]
    yaw = rows[:, 10]
    tab[:, 7] = np.cos(0.5 * yaw)
    tab[:, 10] = np.sin(0.5 * yaw)
    if not np.all(np.diff(tab[:, 0]) > 0):
        raise ValueError("trajectory times must be strictly increasing")
    return tab.astype(np.float32)


def random_lemniscate_table(rng: np.random.Generator, duration: float = 30.0, rate_hz: float = 100.0) -> np.ndarray:
    """One draw of the BASELINE synthetic trajectory family: A~U(1,3) m,
    period~U(6,12) s, phase~U(0,2pi)."""
    A, Tp, ph = rng.uniform(1.0, 3.0), rng.uniform(6.0, 12.0), rng.uniform(0.0, 2.0 * np.pi)
    return csv_rows_to_table(lemniscate(A, Tp, ph, duration=duration, rate_hz=rate_hz))


def interp_table(table: np.ndarray, t: np.ndarray) -> np.ndarray:
    """NumPy statement of the table interpolation (linear per column, clamped,
    quaternion renormalised) used to build explicit reference windows on the host."""
    table = np.asarray(table, np.float32)
    t = np.atleast_1d(np.asarray(t, np.float32))
    tt = table[:, 0]
    idx = np.clip(np.searchsorted(tt, t, side="right") - 1, 0, len(tt) - 2)
    a, b = table[idx], table[idx + 1]
    al = np.clip((t - a[:, 0]) / (b[:, 0] - a[:, 0]), 0.0, 1.0).astype(np.float32)[:, None]
    out = a[:, 1:] + al * (b[:, 1:] - a[:, 1:])
    out[:, 6:10] /= np.linalg.norm(out[:, 6:10], axis=1, keepdims=True)
    return out.astype(np.float32)
