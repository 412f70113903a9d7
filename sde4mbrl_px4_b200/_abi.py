"""ctypes mirror of ``include/sdempc.h`` (struct layouts, constants, loader).

The C header is the source of truth; ``tests/test_abi.py`` checks that the
struct sizes here match ``sizeof`` as compiled by gcc and that the shared
library exports every symbol the header declares.
"""
from __future__ import annotations

import ctypes as C
import os

NX = 13
MAX_NU = 8
MAX_H = 32
NNOISE = 6
TRACE_W = 8
MODEL_MAGIC = 0x4D454453
MODEL_VERSION = 1

F_FRAME_ENU = 1
F_NO_SHIFT = 2
F_SPECULATIVE_LS = 4
F_SEQUENTIAL_LS = 8
F_GROUP = 16
F_NO_CLUSTER = 32
F_TENSOR = 64

OK, EINVAL, ECUDA, ENOMEM, ESTATE = 0, -1, -2, -3, -4

_f = C.c_float
_i = C.c_int32


class Config(C.Structure):
    """``sdempc_config`` — flattened YAML schema (launch/iris_sitl_traj_mpc.yaml:1-85)."""

    _fields_ = [
        ("nu", _i), ("horizon", _i), ("num_particles", _i), ("max_iter", _i),
        ("max_no_improvement_iter", _i), ("maxls", _i), ("reset_option", _i), ("flags", C.c_uint32),
        ("dt", _f * MAX_H), ("discount", _f),
        ("u_lo", _f * MAX_NU), ("u_hi", _f * MAX_NU), ("uref", _f * MAX_NU), ("uerr", _f),
        ("perr", _f * 3), ("verr", _f * 3), ("qerr", _f * 3), ("werr", _f * 3),
        ("res_mult", _f), ("u_slew_coeff", _f),
        ("init_stepsize", _f), ("max_stepsize", _f), ("coef", _f), ("decrease_factor", _f),
        ("increase_factor", _f), ("atol", _f), ("rtol", _f), ("beta_init", _f),
        ("u_slew_constr_coeff", _f), ("u_slew_lo", _f * MAX_NU), ("u_slew_hi", _f * MAX_NU),
        ("moment_scale", _f),
    ]


class ModelHeader(C.Structure):
    """``sdempc_model_header``."""

    _fields_ = [
        ("magic", C.c_uint32), ("version", C.c_uint32),
        ("nu", _i), ("n_in", _i), ("width", _i), ("n_hidden", _i), ("n_out", _i),
        ("mass", _f), ("gravity", _f), ("k_thrust", _f), ("inertia", _f * 3),
        ("mixer", _f * (3 * MAX_NU)), ("sigma_prior", _f * NNOISE),
    ]


class Info(C.Structure):
    """``sdempc_info`` — the opt_state scalars the node reads (sde_control.py:444-450)."""

    _fields_ = [
        ("avg_linesearch", _f), ("stepsize", _f), ("num_steps", _f), ("grad_sqr", _f),
        ("avg_stepsize", _f), ("init_cost", _f), ("opt_cost", _f), ("solve_time_us", _f),
    ]


INFO_FIELDS = [n for n, _ in Info._fields_]

_fp = C.POINTER(C.c_float)
_u64p = C.POINTER(C.c_uint64)


class SolveArgs(C.Structure):
    """``sdempc_solve_args``."""

    _fields_ = [
        ("B", _i), ("x", _fp), ("curr_t", _fp), ("xdes", _fp), ("xref_win", _fp),
        ("rng", _u64p), ("u_plan", _fp), ("x_evol", _fp), ("info", C.POINTER(Info)),
        ("xi_override", _fp), ("trace", _fp),
    ]


HEADER_SYMBOLS = [
    "sdempc_create", "sdempc_set_trajectory", "sdempc_state_from_traj", "sdempc_reset",
    "sdempc_solve_ex", "sdempc_solve", "sdempc_rollout", "sdempc_closed_loop", "sdempc_stage",
    "sdempc_launch_timed", "sdempc_sync", "sdempc_device_out", "sdempc_fetch", "sdempc_fetch_direct", "sdempc_host_register", "sdempc_host_unregister", "sdempc_launch_count", "sdempc_last_launch_ms", "sdempc_kernel_info", "sdempc_probe_fp32",
    "sdempc_destroy", "sdempc_last_error", "sdempc_version",
]

REPO_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libsdempc.so")

_lib = None


def load_library(path: str | None = None) -> C.CDLL:
    """Load ``libsdempc.so`` (built in-tree by ``__graft_entry__.build()``).

    There is no fallback: a missing library is an error, never a silent CPU path.
    Loading the library does not create a CUDA context (fork rule).
    """
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or os.environ.get("SDEMPC_LIB") or LIB_PATH   # SDEMPC_LIB: experiment builds (tools/ab_variant.sh)
    if not os.path.exists(p):
        raise RuntimeError(
            f"{p} not found: build the CUDA extension first (python -c 'import __graft_entry__ as g; g.build()'). "
            "There is no CPU fallback for the MPC solve."
        )
    lib = C.CDLL(p)
    vp = C.c_void_p
    lib.sdempc_create.argtypes = [C.POINTER(Config), C.c_void_p, C.c_size_t, C.c_int, C.POINTER(vp)]
    lib.sdempc_set_trajectory.argtypes = [vp, _fp, C.c_int]
    lib.sdempc_state_from_traj.argtypes = [vp, _fp, C.c_int, _fp]
    lib.sdempc_reset.argtypes = [vp, C.c_int, _fp, _fp, _fp, C.POINTER(Info)]
    lib.sdempc_solve_ex.argtypes = [vp, C.POINTER(SolveArgs)]
    lib.sdempc_solve.argtypes = [vp, C.c_int, _fp, _fp, _fp, _u64p, _fp, _fp, C.POINTER(Info), _fp]
    lib.sdempc_rollout.argtypes = [vp, C.c_int, _fp, _fp, _fp, _fp, _u64p, _fp, _fp, _fp, _fp, _fp, _fp]
    lib.sdempc_closed_loop.argtypes = [vp, C.c_int, C.c_int, _fp, _fp, _u64p, _fp, _fp, _fp]
    lib.sdempc_stage.argtypes = [vp, C.POINTER(SolveArgs)]
    lib.sdempc_launch_timed.argtypes = [vp, C.c_int, C.c_int, _fp]
    lib.sdempc_sync.argtypes = [vp]
    lib.sdempc_device_out.argtypes = [vp, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
    lib.sdempc_fetch.argtypes = [vp, C.POINTER(SolveArgs)]
    lib.sdempc_fetch_direct.argtypes = [vp, C.POINTER(SolveArgs)]
    lib.sdempc_host_register.argtypes = [C.c_void_p, C.c_size_t]
    lib.sdempc_host_unregister.argtypes = [C.c_void_p]
    lib.sdempc_launch_count.argtypes = [vp]
    lib.sdempc_launch_count.restype = C.c_int64
    lib.sdempc_last_launch_ms.argtypes = [vp]
    lib.sdempc_last_launch_ms.restype = C.c_float
    lib.sdempc_kernel_info.argtypes = [vp, C.POINTER(_i * 6)]
    lib.sdempc_probe_fp32.argtypes = [C.c_int, _fp]
    lib.sdempc_destroy.argtypes = [vp]
    lib.sdempc_destroy.restype = None
    lib.sdempc_last_error.restype = C.c_char_p
    lib.sdempc_version.restype = C.c_char_p
    if path is None:
        _lib = lib
    return lib
