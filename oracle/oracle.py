"""ctypes binding of the CPU oracle (``oracle/sdempc_oracle.c``).

TEST INFRASTRUCTURE ONLY — see the header of ``sdempc_oracle_impl.h``.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may
import this module.  PARITY UNPINNED: the reference's solver source is absent
(un-vendored JAX package ``sde4mbrl``); this restates SURVEY.md section 8(a) [SPEC].
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from sde4mbrl_px4_b200 import _abi

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libsdempc_oracle.so")

_lib = None


def build(force: bool = False) -> str:
    src = [os.path.join(HERE, f) for f in ("sdempc_oracle.c", "sdempc_oracle_impl.h", "det_math.h")] + [
        os.path.join(HERE, "..", "include", "sdempc.h")]
    if force or not os.path.exists(LIB) or (
            all(os.path.exists(s) for s in src) and os.path.getmtime(LIB) < max(os.path.getmtime(s) for s in src)):
        subprocess.check_call(["make", "-C", HERE, "-s"] + (["-B"] if force else []))
    return LIB


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        _lib = C.CDLL(LIB)
    return _lib


def _ptr(a, ct):
    return None if a is None else a.ctypes.data_as(C.POINTER(ct))


class Oracle:
    """CPU reference solve; ``dtype`` 'f32' (parity oracle / timed baseline) or 'f64' (gradient checks)."""

    def __init__(self, cfg: _abi.Config, blob: bytes, dtype: str = "f32"):
        assert dtype in ("f32", "f64")
        self.cfg, self.dtype = cfg, dtype
        self.np = np.float32 if dtype == "f32" else np.float64
        self.ct = C.c_float if dtype == "f32" else C.c_double
        self.H, self.nu, self.P = cfg.horizon, cfg.nu, cfg.num_particles
        self._l = lib()
        self._h = C.c_void_p()
        buf = (C.c_char * len(blob)).from_buffer_copy(blob)
        rc = self._fn("create", C.c_int)(C.byref(cfg), buf, C.c_size_t(len(blob)), C.byref(self._h))
        if rc:
            raise RuntimeError(f"oracle create failed: {rc}")

    def _fn(self, name, restype=C.c_int):
        f = getattr(self._l, f"oracle_{self.dtype}_{name}")
        f.restype = restype
        return f

    def __del__(self):
        try:
            if self._h:
                self._fn("destroy", None)(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    def _a(self, a, shape=None):
        if a is None:
            return None
        a = np.ascontiguousarray(a, dtype=self.np)
        if shape is not None:
            a = a.reshape(shape)
        return a

    def set_trajectory(self, table):
        t = np.ascontiguousarray(table, np.float32)
        rc = self._fn("set_trajectory")(self._h, _ptr(t, C.c_float), C.c_int(t.shape[0]))
        if rc:
            raise RuntimeError(f"oracle set_trajectory failed: {rc}")
        self.T = t.shape[0]

    def traj_internal(self):
        p = self._fn("traj_internal", C.POINTER(C.c_float))(self._h)
        return np.ctypeslib.as_array(p, shape=(self.T, 14)).copy()

    def state_from_traj(self, t):
        t = self._a(np.atleast_1d(t))
        out = np.zeros((t.shape[0], 13), self.np)
        rc = self._fn("state_from_traj")(self._h, _ptr(t, self.ct), C.c_int(t.shape[0]), _ptr(out, self.ct))
        if rc:
            raise RuntimeError(f"oracle state_from_traj failed: {rc}")
        return out

    def enu2ned(self, x):
        x = self._a(x, (-1, 13))
        o = np.zeros_like(x)
        self._fn("enu2ned", None)(_ptr(x, self.ct), _ptr(o, self.ct), C.c_int(x.shape[0]))
        return o

    def noise(self, seed, tick, P=None, H=None, sub0=0):
        P, H = P or self.P, H or self.H
        xi = np.zeros((P, H, 6), self.np)
        self._fn("noise", None)(C.c_uint64(seed), C.c_uint64(tick), C.c_int(P), C.c_int(H), C.c_int(sub0), _ptr(xi, self.ct))
        return xi

    def reset(self, B):
        u = np.zeros((B, self.H, self.nu), self.np)
        info = np.zeros((B, 8), self.np)
        self._fn("reset")(self._h, C.c_int(B), _ptr(u, self.ct), _ptr(info, self.ct))
        return u, info

    def rollout(self, x, u, u_prev, curr_t=None, xdes=None, xref_win=None, rng=None, xi=None, want_grad=True):
        x = self._a(x, (-1, 13))
        B = x.shape[0]
        u = self._a(u, (B, self.H, self.nu))
        u_prev = self._a(u_prev, (B, self.nu))
        curr_t, xdes = self._a(curr_t, (B,)), self._a(xdes, (B, 13))
        xref_win = self._a(xref_win, (B, self.H + 1, 13))
        xi = self._a(xi, (B, self.P, self.H, 6))
        rng = None if rng is None else np.ascontiguousarray(rng, np.uint64).reshape(B, 2)
        cost = np.zeros((B,), self.np)
        grad = np.zeros((B, self.H, self.nu), self.np) if want_grad else None
        xe = np.zeros((B, self.H + 1, 13), self.np)
        rc = self._fn("rollout_batch")(
            self._h, C.c_int(B), _ptr(x, self.ct), _ptr(curr_t, self.ct), _ptr(xdes, self.ct), _ptr(xref_win, self.ct),
            _ptr(rng, C.c_uint64), _ptr(xi, self.ct), _ptr(u, self.ct), _ptr(u_prev, self.ct),
            _ptr(cost, self.ct), _ptr(grad, self.ct), _ptr(xe, self.ct))
        if rc:
            raise RuntimeError(f"oracle rollout failed: {rc}")
        return cost, grad, xe

    def solve(self, x, u_plan, info, curr_t=None, xdes=None, xref_win=None, rng=None, xi=None, want_trace=False):
        """Returns (u_plan', x_evol, info', trace|None); inputs are not modified."""
        x = self._a(x, (-1, 13))
        B = x.shape[0]
        u = self._a(u_plan, (B, self.H, self.nu)).copy()
        info = self._a(info, (B, 8)).copy()
        curr_t, xdes = self._a(curr_t, (B,)), self._a(xdes, (B, 13))
        xref_win = self._a(xref_win, (B, self.H + 1, 13))
        xi = self._a(xi, (B, self.P, self.H, 6))
        rng = None if rng is None else np.ascontiguousarray(rng, np.uint64).reshape(B, 2)
        xe = np.zeros((B, self.H + 1, 13), self.np)
        trace = np.zeros((B, self.cfg.max_iter, _abi.TRACE_W), self.np) if want_trace else None
        rc = self._fn("solve")(
            self._h, C.c_int(B), _ptr(x, self.ct), _ptr(curr_t, self.ct), _ptr(xdes, self.ct), _ptr(xref_win, self.ct),
            _ptr(rng, C.c_uint64), _ptr(u, self.ct), _ptr(xe, self.ct), _ptr(info, self.ct), _ptr(xi, self.ct),
            _ptr(trace, self.ct))
        if rc:
            raise RuntimeError(f"oracle solve failed: {rc}")
        return u, xe, info, trace

    def solve_forced(self, x, u_plan, info, forced_trace, forced_iters, curr_t=None, xdes=None, xref_win=None, rng=None, xi=None):
        """Teacher-forced replay: the trial counts, accept / reject decisions and iteration counts come from
        ``forced_trace[B, max_iter, 8]`` / ``forced_iters[B]`` (another solver's decision trace); returns
        (u_plan', x_evol, info', trace, own) with ``own[B, max_iter, 4]`` = (this oracle's own trial count, own
        accept, Armijo margin of the last forced trial, accept margin J_x - J_trial) per iteration."""
        x = self._a(x, (-1, 13))
        B = x.shape[0]
        u = self._a(u_plan, (B, self.H, self.nu)).copy()
        info = self._a(info, (B, 8)).copy()
        curr_t, xdes = self._a(curr_t, (B,)), self._a(xdes, (B, 13))
        xref_win = self._a(xref_win, (B, self.H + 1, 13))
        xi = self._a(xi, (B, self.P, self.H, 6))
        rng = None if rng is None else np.ascontiguousarray(rng, np.uint64).reshape(B, 2)
        forced = self._a(forced_trace, (B, self.cfg.max_iter, _abi.TRACE_W))
        fit = self._a(forced_iters, (B,))
        xe = np.zeros((B, self.H + 1, 13), self.np)
        trace = np.zeros((B, self.cfg.max_iter, _abi.TRACE_W), self.np)
        own = np.zeros((B, self.cfg.max_iter, 4), self.np)
        rc = self._fn("solve_forced")(
            self._h, C.c_int(B), _ptr(x, self.ct), _ptr(curr_t, self.ct), _ptr(xdes, self.ct), _ptr(xref_win, self.ct),
            _ptr(rng, C.c_uint64), _ptr(u, self.ct), _ptr(xe, self.ct), _ptr(info, self.ct), _ptr(xi, self.ct),
            _ptr(trace, self.ct), _ptr(forced, self.ct), _ptr(fit, self.ct), _ptr(own, self.ct))
        if rc:
            raise RuntimeError(f"oracle solve_forced failed: {rc}")
        return u, xe, info, trace, own

    def closed_loop(self, x0, t0, rng, ticks, want_hist=True):
        x0 = self._a(x0, (-1, 13))
        R = x0.shape[0]
        t0 = self._a(t0, (R,))
        rng = np.ascontiguousarray(rng, np.uint64).reshape(R, 2)
        xh = np.zeros((R, ticks + 1, 13), self.np) if want_hist else None
        uh = np.zeros((R, ticks, self.nu), self.np) if want_hist else None
        stats = np.zeros((R, 4), self.np)
        rc = self._fn("closed_loop")(self._h, C.c_int(R), C.c_int(ticks), _ptr(x0, self.ct), _ptr(t0, self.ct),
                                     _ptr(rng, C.c_uint64), _ptr(xh, self.ct), _ptr(uh, self.ct), _ptr(stats, self.ct))
        if rc:
            raise RuntimeError(f"oracle closed_loop failed: {rc}")
        return xh, uh, stats


def set_threads(n: int) -> int:
    """Number of OpenMP threads for the batched oracle calls (torchrun exports OMP_NUM_THREADS=1)."""
    lib().oracle_set_threads(C.c_int(int(n)))
    return int(lib().oracle_get_max_threads())


def philox4x32_10(ctr, key):
    ctr = np.ascontiguousarray(ctr, np.uint32)
    key = np.ascontiguousarray(key, np.uint32)
    out = np.zeros(4, np.uint32)
    f = lib().oracle_philox
    f.restype = None
    f(_ptr(ctr, C.c_uint32), _ptr(key, C.c_uint32), _ptr(out, C.c_uint32))
    return out
