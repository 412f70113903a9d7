"""Thin NumPy front end of the C ABI (``include/sdempc.h``).

One ``MPCSolver`` owns one ``sdempc_t`` handle.  Construction is host-only; the
CUDA context is created by the first compute call in the calling process, which
is what lets the reference node fork its solver process after building the
solver objects (sde_control.py:66-75, 723-728).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _abi


def _fp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_float))


def _f32(a, shape=None):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a.reshape(shape) if shape is not None else a


def probe_fp32_tflops(device: int = 0) -> float:
    """FP32 FMA-pipe throughput of ``device`` measured now (``sdempc_probe_fp32``): the roofline denominator."""
    v = C.c_float(0.0)
    lib = _abi.load_library()
    rc = lib.sdempc_probe_fp32(device, C.byref(v))
    if rc != 0:
        raise RuntimeError(f"sdempc error {rc}: {lib.sdempc_last_error().decode()}")
    return float(v.value)


class MPCSolver:
    def __init__(self, cfg: _abi.Config, model_blob: bytes, device: int = 0, lib_path: str | None = None):
        self.lib = _abi.load_library(lib_path)
        self.cfg = cfg
        self.H, self.nu, self.P = cfg.horizon, cfg.nu, cfg.num_particles
        self._h = C.c_void_p()
        buf = (C.c_char * len(model_blob)).from_buffer_copy(model_blob)
        self._check(self.lib.sdempc_create(C.byref(cfg), buf, len(model_blob), device, C.byref(self._h)))
        self.has_trajectory = False
        self._staged = None
        self._pinned = []

    # ------------------------------------------------------------------ helpers
    def _check(self, rc: int):
        if rc != 0:
            raise RuntimeError(f"sdempc error {rc}: {self.lib.sdempc_last_error().decode()}")

    def close(self):
        for a in getattr(self, "_pinned", []):
            self.lib.sdempc_host_unregister(a.ctypes.data)
        self._pinned = []
        if getattr(self, "_h", None):
            self.lib.sdempc_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ trajectory
    def set_trajectory(self, table):
        t = _f32(table)
        assert t.ndim == 2 and t.shape[1] == 14
        self._check(self.lib.sdempc_set_trajectory(self._h, _fp(t), t.shape[0]))
        self.has_trajectory = True

    def state_from_traj(self, t):
        t = _f32(np.atleast_1d(t))
        out = np.zeros((t.shape[0], 13), np.float32)
        self._check(self.lib.sdempc_state_from_traj(self._h, _fp(t), t.shape[0], _fp(out)))
        return out

    # ------------------------------------------------------------------ reset / solve
    def reset(self, B: int = 1):
        u = np.zeros((B, self.H, self.nu), np.float32)
        info = (_abi.Info * B)()
        self._check(self.lib.sdempc_reset(self._h, B, None, None, _fp(u), info))
        return u, np.frombuffer(info, dtype=np.float32).reshape(B, 8).copy()

    def _args(self, x, u_plan, info, curr_t, xdes, xref_win, rng, xi, want_trace):
        x = _f32(x, (-1, 13))
        B = x.shape[0]
        keep = dict(
            x=x, u=_f32(u_plan, (B, self.H, self.nu)).copy(), info=_f32(info, (B, 8)).copy(),
            curr_t=_f32(curr_t, (B,)), xdes=_f32(xdes, (B, 13)), xref=_f32(xref_win, (B, self.H + 1, 13)),
            xi=_f32(xi, (B, self.P, self.H, 6)),
            rng=None if rng is None else np.ascontiguousarray(rng, np.uint64).reshape(B, 2),
            xe=np.zeros((B, self.H + 1, 13), np.float32),
            trace=np.zeros((B, self.cfg.max_iter, _abi.TRACE_W), np.float32) if want_trace else None,
        )
        a = _abi.SolveArgs()
        a.B = B
        a.x, a.curr_t, a.xdes, a.xref_win = _fp(keep["x"]), _fp(keep["curr_t"]), _fp(keep["xdes"]), _fp(keep["xref"])
        a.rng = None if keep["rng"] is None else keep["rng"].ctypes.data_as(C.POINTER(C.c_uint64))
        a.u_plan, a.x_evol = _fp(keep["u"]), _fp(keep["xe"])
        a.info = keep["info"].ctypes.data_as(C.POINTER(_abi.Info))
        a.xi_override, a.trace = _fp(keep["xi"]), _fp(keep["trace"])
        return a, keep

    def solve(self, x, u_plan, info, curr_t=None, xdes=None, xref_win=None, rng=None, xi=None, want_trace=False):
        """Batched ``m_mpc``: returns (u_plan', x_evol, info', trace|None); inputs are not modified."""
        a, keep = self._args(x, u_plan, info, curr_t, xdes, xref_win, rng, xi, want_trace)
        self._check(self.lib.sdempc_solve_ex(self._h, C.byref(a)))
        return keep["u"], keep["xe"], keep["info"], keep["trace"]

    # ------------------------------------------------------------------ zero-copy batches
    def pin(self, *arrays) -> bool:
        """Page-lock caller arrays in place (sdempc_host_register): ``solve`` / ``stage`` then send them by DMA straight from
        their memory instead of through the staging block.  Released by ``close``.  False if any registration failed."""
        ok = True
        for a in arrays:
            if a is None:
                continue
            if self.lib.sdempc_host_register(a.ctypes.data, a.nbytes) == 0:
                self._pinned.append(a)
            else:
                ok = False
        return ok

    def pinned_batch(self, B: int, window: bool = True) -> "PinnedBatch":
        """Page-locked buffers for ``solve_pinned``: the library then copies straight between them and the device, with no
        staging copy on the host.  ``window``: explicit reference windows (``xref``) instead of trajectory times (``curr_t``)."""
        return PinnedBatch(self, B, window)

    def solve_pinned(self, pb: "PinnedBatch"):
        """``m_mpc`` on a PinnedBatch, in place: reads pb.x, pb.u (plan, shifted by the solve), pb.info, pb.xref | pb.curr_t,
        pb.rng; writes pb.u, pb.xe, pb.info.  Same call (sdempc_solve_ex) and same results as ``solve``."""
        self._check(self.lib.sdempc_solve_ex(self._h, C.byref(pb.args)))
        return pb.u, pb.xe, pb.info

    def rollout(self, x, u, u_prev, curr_t=None, xdes=None, xref_win=None, rng=None, xi=None, want_grad=True):
        """value_and_grad of the MPC objective at ``u``: (cost[B], grad[B,H,nu]|None, x_evol[B,H+1,13])."""
        x = _f32(x, (-1, 13))
        B = x.shape[0]
        u = _f32(u, (B, self.H, self.nu))
        u_prev = _f32(u_prev, (B, self.nu))
        curr_t, xdes, xref = _f32(curr_t, (B,)), _f32(xdes, (B, 13)), _f32(xref_win, (B, self.H + 1, 13))
        xi = _f32(xi, (B, self.P, self.H, 6))
        rng = None if rng is None else np.ascontiguousarray(rng, np.uint64).reshape(B, 2)
        cost = np.zeros((B,), np.float32)
        grad = np.zeros((B, self.H, self.nu), np.float32) if want_grad else None
        xe = np.zeros((B, self.H + 1, 13), np.float32)
        self._check(self.lib.sdempc_rollout(
            self._h, B, _fp(x), _fp(curr_t), _fp(xdes), _fp(xref),
            None if rng is None else rng.ctypes.data_as(C.POINTER(C.c_uint64)), _fp(xi), _fp(u), _fp(u_prev),
            _fp(cost), _fp(grad), _fp(xe)))
        return cost, grad, xe

    def closed_loop(self, x0, t0, rng, ticks: int, want_hist: bool = True):
        x0 = _f32(x0, (-1, 13))
        R = x0.shape[0]
        t0 = _f32(t0, (R,))
        rng = np.ascontiguousarray(rng, np.uint64).reshape(R, 2)
        xh = np.zeros((R, ticks + 1, 13), np.float32) if want_hist else None
        uh = np.zeros((R, ticks, self.nu), np.float32) if want_hist else None
        stats = np.zeros((R, 4), np.float32)
        self._check(self.lib.sdempc_closed_loop(self._h, R, ticks, _fp(x0), _fp(t0),
                                                rng.ctypes.data_as(C.POINTER(C.c_uint64)), _fp(xh), _fp(uh), _fp(stats)))
        return xh, uh, stats

    # ------------------------------------------------------------------ staged API (measurement)
    def stage(self, x, u_plan, info, curr_t=None, xdes=None, xref_win=None, rng=None, xi=None):
        a, keep = self._args(x, u_plan, info, curr_t, xdes, xref_win, rng, xi, False)
        self._check(self.lib.sdempc_stage(self._h, C.byref(a)))
        self._staged = (a, keep)

    def launch_timed(self, n: int, flush_l2: bool = True) -> np.ndarray:
        ms = np.zeros((n,), np.float32)
        self._check(self.lib.sdempc_launch_timed(self._h, n, 1 if flush_l2 else 0, _fp(ms)))
        return ms

    def fetch(self):
        a, keep = self._staged
        self._check(self.lib.sdempc_fetch(self._h, C.byref(a)))
        return keep["u"], keep["xe"], keep["info"]

    def fetch_into(self, u, xe, info, direct: bool = False):
        """D2H of the staged solve's outputs into caller-provided float32 arrays (e.g. slices of a shared-memory result
        buffer): u[B,H,nu], xe[B,H+1,13], info[B,8].  ``direct=True`` skips the pinned staging copy (`sdempc_fetch_direct`):
        for destinations the caller has page-locked with cudaHostRegister."""
        a0, keep = self._staged
        B = keep["x"].shape[0]
        for arr, shape in ((u, (B, self.H, self.nu)), (xe, (B, self.H + 1, 13)), (info, (B, 8))):
            if arr.dtype != np.float32 or tuple(arr.shape) != shape or not arr.flags["C_CONTIGUOUS"]:
                raise ValueError(f"fetch_into: expected a C-contiguous float32 array of shape {shape}")
        a = _abi.SolveArgs()
        C.memmove(C.byref(a), C.byref(a0), C.sizeof(a))
        a.u_plan, a.x_evol, a.info, a.trace = _fp(u), _fp(xe), info.ctypes.data_as(C.POINTER(_abi.Info)), None
        self._check((self.lib.sdempc_fetch_direct if direct else self.lib.sdempc_fetch)(self._h, C.byref(a)))

    def sync(self):
        self._check(self.lib.sdempc_sync(self._h))

    def device_out(self):
        """(device pointer, nbytes, layout) of the OUT block of the staged solve; layout maps
        name -> (byte offset, shape) of x_evol / u / info inside the block."""
        ptr, n = C.c_void_p(), C.c_size_t()
        self._check(self.lib.sdempc_device_out(self._h, C.byref(ptr), C.byref(n)))
        B = self._staged[1]["x"].shape[0]
        a16 = lambda b: (b + 15) & ~15
        o_p = a16(B * (self.H + 1) * 13 * 4)
        o_i = o_p + a16(B * self.H * self.nu * 4)
        layout = {"x_evol": (0, (B, self.H + 1, 13)), "u": (o_p, (B, self.H, self.nu)), "info": (o_i, (B, 8))}
        return int(ptr.value), int(n.value), layout

    def last_launch_ms(self) -> float:
        return float(self.lib.sdempc_last_launch_ms(self._h))

    def launch_count(self) -> int:
        return int(self.lib.sdempc_launch_count(self._h))

    def kernel_info(self) -> dict:
        out = (C.c_int32 * 6)()
        self._check(self.lib.sdempc_kernel_info(self._h, C.byref(out)))
        keys = ("threads_per_cta", "smem_bytes", "problems_per_cta", "regs_per_thread", "ctas", "sm_count")
        return dict(zip(keys, [int(v) for v in out]))


class PinnedBatch:
    """Caller-side buffers of one batched solve, page-locked through sdempc_host_register (plain pageable arrays if the
    registration fails: the call still works, through the staging copies)."""

    def __init__(self, solver: MPCSolver, B: int, window: bool = True):
        self._lib = solver.lib
        H, nu = solver.H, solver.nu
        self.x = np.zeros((B, 13), np.float32)
        self.u = np.zeros((B, H, nu), np.float32)
        self.info = np.zeros((B, 8), np.float32)
        self.xref = np.zeros((B, H + 1, 13), np.float32) if window else None
        self.curr_t = None if window else np.zeros((B,), np.float32)
        self.rng = np.zeros((B, 2), np.uint64)
        self.xe = np.zeros((B, H + 1, 13), np.float32)
        self._registered = []
        for a in (self.x, self.u, self.info, self.xref, self.curr_t, self.rng, self.xe):
            if a is not None and self._lib.sdempc_host_register(a.ctypes.data, a.nbytes) == 0:
                self._registered.append(a)
        self.pinned = len(self._registered) == 6
        a = _abi.SolveArgs()
        a.B = B
        a.x, a.curr_t, a.xdes, a.xref_win = _fp(self.x), _fp(self.curr_t), None, _fp(self.xref)
        a.rng = self.rng.ctypes.data_as(C.POINTER(C.c_uint64))
        a.u_plan, a.x_evol = _fp(self.u), _fp(self.xe)
        a.info = self.info.ctypes.data_as(C.POINTER(_abi.Info))
        a.xi_override, a.trace = None, None
        self.args = a

    def close(self):
        for a in self._registered:
            self._lib.sdempc_host_unregister(a.ctypes.data)
        self._registered = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
