"""``load_mpc_from_cfgfile`` — the controller interface the reference node imports
(``from sde4mbrlExamples.rotor_uav.sde_mpc_design import load_mpc_from_cfgfile``,
sde_control.py:12) re-exposed over the sm_100a CUDA library.

    cfg_dict, (m_reset, m_mpc), state_from_traj, aux = load_mpc_from_cfgfile(path, convert_to_enu=True)

mirrors sde_control.py:685; ``m_reset(x=, rng=, xdes=)`` mirrors :345-346/:702 and
``m_mpc(x, rng, opt_state, curr_t=, xdes=)`` mirrors :400-416/:713-719.  Building the
callables is host-only: the CUDA context is created by the first ``m_mpc`` call in
the calling process (the node forks its solver process after loading, :723-728).
There is no CPU fallback.
"""
from __future__ import annotations

import os
from typing import NamedTuple

import numpy as np

from . import _abi, config, model_io, solver, trajectory
from .utils import PRNGKey, enu2ned, ned2enu, split  # noqa: F401  (re-exported)


class HostArray(np.ndarray):
    """NumPy array with JAX's ``.block_until_ready()`` (sde_control.py:420, 707, 718).
    Results are already synchronised in host memory when a solve returns."""

    def block_until_ready(self):
        return self


def _wrap(a) -> HostArray:
    return np.asarray(a).view(HostArray)


class OptState(NamedTuple):
    """Optimiser state carried from tick to tick; the scalar fields are the ones the node
    copies into ``OptMPCState`` (sde_control.py:444-450; msg/OptMPCState.msg:5-24)."""

    yk: HostArray            # current plan [H, nu] (the next solve warm-starts from it)
    avg_linesearch: float
    stepsize: float
    num_steps: float
    grad_sqr: float
    avg_stepsize: float
    init_cost: float
    opt_cost: float
    solve_time_us: float = 0.0

    def info_array(self) -> np.ndarray:
        return np.asarray([self.avg_linesearch, self.stepsize, self.num_steps, self.grad_sqr,
                           self.avg_stepsize, self.init_cost, self.opt_cost, self.solve_time_us], np.float32)


def _resolve_model(spec: str) -> model_io.SDEModel:
    if spec.startswith("synthetic:"):
        return model_io.synthetic_model(spec.split(":", 1)[1])
    path = os.path.expanduser(spec)
    if not os.path.exists(path):
        raise FileNotFoundError(
            f"learned_model_params: {spec!r} not found. Point it at an .npz model of sde4mbrl_px4_b200.model_io "
            "or use 'synthetic:iris' / 'synthetic:hexa'.")
    return model_io.SDEModel.load(path)


def _resolve_trajectory(spec: str) -> np.ndarray:
    if spec.startswith("synthetic:"):
        kind = spec.split(":", 1)[1]
        if kind != "lemniscate":
            raise ValueError(f"unknown synthetic trajectory {kind!r}")
        return trajectory.csv_rows_to_table(trajectory.lemniscate())
    return trajectory.csv_rows_to_table(trajectory.load_csv(spec))


class MPCController:
    """One loaded controller (trajectory tracker or set-point controller)."""

    def __init__(self, cfg_path_or_dict, convert_to_enu: bool = True, device: int = 0, **overrides):
        cfgd = config.load_yaml(cfg_path_or_dict) if isinstance(cfg_path_or_dict, str) else dict(cfg_path_or_dict)
        self.cfg_dict = cfgd
        self.model = _resolve_model(str(cfgd["learned_model_params"]))
        self.cfg = config.build_config(cfgd, convert_to_enu=convert_to_enu, **overrides)
        if self.cfg.nu != self.model.nu:
            raise config.ConfigError(f"config has {self.cfg.nu} inputs but the model has {self.model.nu}")
        self.solver = solver.MPCSolver(self.cfg, self.model.to_blob(), device=device)
        self.has_trajectory = "trajectory_path" in cfgd and cfgd["trajectory_path"] is not None
        if self.has_trajectory:
            self.table = _resolve_trajectory(str(cfgd["trajectory_path"]))
            self.solver.set_trajectory(self.table)
        self.cfg_dict["_time_steps"] = config.time_steps(cfgd)
        self.H, self.nu = self.cfg.horizon, self.cfg.nu

    # --- the callables handed to the node ------------------------------------------
    def state_from_traj(self, t):
        """13-state (external frame) of the reference trajectory at time ``t`` (sde_control.py:206)."""
        return _wrap(self.solver.state_from_traj(np.float32(t))[0])

    def m_reset(self, x=None, rng=None, xdes=None) -> OptState:
        u, info = self.solver.reset(1)
        i = info[0]
        return OptState(_wrap(u[0]), *[float(v) for v in i[:7]], 0.0)

    def m_mpc(self, x, rng, opt_state: OptState, curr_t=0.0, xdes=None, shift: int = 1):
        """One MPC solve.  Trajectory controllers track ``state_from_traj(curr_t + ...)`` and ignore
        ``xdes`` ("xdes does not matter here", sde_control.py:408); set-point controllers track ``xdes``.

        Warm start: the solve shifts ``opt_state.yk`` left by one step (last row repeated) and uses its
        first row as the control being applied (slew reference), i.e. it assumes ONE solve per ``dt[0]``
        — the rate the node runs at (``_dt_usec_*`` = ``_time_steps[0]``, sde_control.py:167-177).  A
        caller that solves every n-th step passes ``shift=n``: the plan is advanced n - 1 further rows
        here, so the first row handed to the library is the control applied during the last elapsed step."""
        x = np.asarray(x, np.float32).reshape(1, 13)
        rng = np.asarray(rng, np.uint64).reshape(1, 2)
        plan = np.asarray(opt_state.yk, np.float32)
        if shift < 1:
            raise ValueError("shift must be >= 1 (SDEMPC_F_NO_SHIFT / no_shift=True disables the shift for a handle)")
        if shift > 1:
            k = min(shift - 1, self.H - 1)
            plan = np.concatenate([plan[k:], np.repeat(plan[-1:], k, axis=0)], axis=0)
        kw = dict(curr_t=np.asarray([curr_t], np.float32)) if self.has_trajectory else dict(
            xdes=np.asarray(x if xdes is None else xdes, np.float32).reshape(1, 13))
        u, xe, info, _ = self.solver.solve(x, plan[None], opt_state.info_array()[None], rng=rng, **kw)
        i = info[0]
        new_state = OptState(_wrap(u[0]), *[float(v) for v in i[:8]])
        new_rng = rng[0].copy()
        new_rng[1] += np.uint64(1)
        return _wrap(u[0]), new_state, new_rng, _wrap(xe[0])


def load_mpc_from_cfgfile(path, convert_to_enu: bool = True, device: int = 0, **overrides):
    """Returns ``(cfg_dict, (m_reset, m_mpc), state_from_traj | None, controller)``.

    ``state_from_traj`` is ``None`` when the YAML has no ``trajectory_path`` — the node
    asserts exactly that for its set-point controller (sde_control.py:164, 177)."""
    ctl = MPCController(path, convert_to_enu=convert_to_enu, device=device, **overrides)
    return ctl.cfg_dict, (ctl.m_reset, ctl.m_mpc), (ctl.state_from_traj if ctl.has_trajectory else None), ctl


class _Lowered:
    def __init__(self, f):
        self._f = f

    def compile(self):
        return self._f


class _Jitted:
    """Stand-in for ``jax.jit(f)`` so the node's ``jax.jit(f).lower(*a, **k).compile()`` idiom
    (sde_control.py:694, 702, 713) keeps working: the kernels are compiled ahead of time."""

    def __init__(self, f):
        self._f = f

    def __call__(self, *a, **k):
        return self._f(*a, **k)

    def lower(self, *a, **k):
        return _Lowered(self._f)


def jit(f):
    return _Jitted(f)
