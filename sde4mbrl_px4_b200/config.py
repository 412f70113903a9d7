"""YAML controller configuration -> ``sdempc_config``.

Accepts the reference's controller YAMLs unchanged
(/root/reference launch/{iris_sitl,hexa_sitl,hexa}_{traj,posctrl}_mpc.yaml; schema R2 in
SURVEY.md section 8a).  ``learned_model_params`` may point at this framework's
``.npz`` model format, or name a built-in synthetic model as ``synthetic:iris`` /
``synthetic:hexa`` (no learned-model pickle ships with the reference).
"""
from __future__ import annotations

import os
import warnings

import numpy as np
import yaml

from . import _abi

# keys whose semantics live in the un-vendored upstream package and are not on the
# two BASELINE config styles: parsed, kept in cfg_dict, not applied (DESIGN.md "out of scope").
_UNSUPPORTED_COST_KEYS = ("res_sig",)
_UNSUPPORTED_TOP_KEYS = ("state_constr", "use_sysId_model")


class ConfigError(ValueError):
    pass


def load_yaml(path: str) -> dict:
    with open(os.path.expanduser(path), "r") as f:
        d = yaml.safe_load(f)
    if not isinstance(d, dict):
        raise ConfigError(f"{path}: not a mapping")
    return d


def time_steps(cfg: dict) -> np.ndarray:
    """``cfg_dict['_time_steps']``: short_step_dt for t < num_short_dt, then long_step_dt
    (launch/iris_sitl_traj_mpc.yaml:44-48; consumed at sde_control.py:167)."""
    H = int(cfg["horizon"])
    ns = int(cfg.get("num_short_dt", H))
    sdt = float(cfg["short_step_dt"])
    ldt = float(cfg.get("long_step_dt", sdt))
    return np.asarray([sdt if t < ns else ldt for t in range(H)], np.float32)


def build_config(cfg: dict, convert_to_enu: bool = True, strict: bool = False, **overrides) -> _abi.Config:
    """Flatten a parsed YAML dict into the C struct.  ``overrides`` replaces top-level
    scalars after parsing (e.g. ``num_particles=8`` for the BASELINE hexa config,
    ``max_iter=``, ``rtol=``, ``atol=`` for benchmark mode)."""
    c = _abi.Config()
    ic = cfg["input_constr"]
    ids = list(ic["input_id"])
    nu = len(ids)
    if ids != list(range(nu)) or nu > _abi.MAX_NU:
        raise ConfigError(f"input_constr.input_id must be 0..nu-1 with nu <= {_abi.MAX_NU}, got {ids}")
    H = int(cfg["horizon"])
    if not 1 <= H <= _abi.MAX_H - 1:   # the kernels handle rows 0..H of the window with one lane per row
        raise ConfigError(f"horizon must be in 1..{_abi.MAX_H - 1}")
    c.nu, c.horizon = nu, H
    c.num_particles = int(overrides.get("num_particles", cfg.get("num_particles", 1)))
    apg = cfg["apg_mpc"]
    ls = apg["linesearch"]
    c.max_iter = int(overrides.get("max_iter", apg["max_iter"]))
    c.max_no_improvement_iter = int(overrides.get("max_no_improvement_iter", apg.get("max_no_improvement_iter", c.max_iter)))
    c.maxls = int(ls["maxls"])
    ro = str(ls.get("reset_option", "increase"))
    if ro not in ("increase", "conservative"):
        raise ConfigError(f"linesearch.reset_option must be 'increase' or 'conservative', got {ro!r}")
    c.reset_option = 1 if ro == "increase" else 0
    flags = 0
    if convert_to_enu:
        flags |= _abi.F_FRAME_ENU
    if overrides.get("no_shift", False):
        flags |= _abi.F_NO_SHIFT
    if overrides.get("speculative_ls", False):
        flags |= _abi.F_SPECULATIVE_LS
    if overrides.get("sequential_ls", False):
        flags |= _abi.F_SEQUENTIAL_LS
    if overrides.get("group", False):
        flags |= _abi.F_GROUP
    if overrides.get("no_cluster", False):
        flags |= _abi.F_NO_CLUSTER
    if overrides.get("tensor", False):
        flags |= _abi.F_TENSOR
    c.flags = flags
    for t, v in enumerate(time_steps(cfg)):
        c.dt[t] = float(v)
    c.discount = float(cfg.get("discount", 1.0))
    bounds = ic["input_bound"]
    if len(bounds) != nu:
        raise ConfigError("input_constr.input_bound must have one [lo, hi] pair per input")
    enforce = bool(cfg.get("enforce_ubound", True))   # yaml:14; False: the box is not projected onto (bounds +-FLT_MAX)
    cp = cfg["cost_params"]
    uref = list(cp["uref"])
    if len(uref) != nu:
        raise ConfigError("cost_params.uref must have nu entries")
    for i in range(nu):
        c.u_lo[i], c.u_hi[i], c.uref[i] = float(bounds[i][0]), float(bounds[i][1]), float(uref[i])
        if not enforce:
            c.u_lo[i], c.u_hi[i] = -3.0e38, 3.0e38
    c.uerr = float(cp.get("uerr", 0.0))

    def vec3(key):
        v = cp.get(key, 0.0)
        v = [v] * 3 if np.isscalar(v) else list(v)
        if len(v) != 3:
            raise ConfigError(f"cost_params.{key} must be a scalar or 3 values")
        return [float(a) for a in v]

    for key in ("perr", "verr", "qerr", "werr"):
        for i, v in enumerate(vec3(key)):
            getattr(c, key)[i] = v
    c.res_mult = float(cp.get("res_mult", 0.0))
    c.u_slew_coeff = float(cp.get("u_slew_coeff", 0.0))
    # soft input-rate constraint of the position-control YAML (iris_sitl_posctrl_mpc.yaml:40-41)
    sc = cp.get("u_slew_constr", None)
    c.u_slew_constr_coeff = float(cp.get("u_slew_constr_coeff", 0.0)) if sc is not None else 0.0
    if sc is not None:
        if len(sc) != nu or any(len(b) != 2 for b in sc):
            raise ConfigError("cost_params.u_slew_constr must have one [lo, hi] pair per input")
        for i, (lo, hi) in enumerate(sc):
            if not float(lo) <= float(hi):
                raise ConfigError("cost_params.u_slew_constr: lo must not exceed hi")
            c.u_slew_lo[i], c.u_slew_hi[i] = float(lo), float(hi)
    c.init_stepsize = float(ls.get("init_stepsize", apg.get("stepsize", 1.0)))
    c.max_stepsize = float(ls.get("max_stepsize", 1.0))
    c.coef = float(ls.get("coef", 0.01))
    c.decrease_factor = float(ls.get("decrease_factor", 0.7))
    c.increase_factor = float(ls.get("increase_factor", 1.3))
    c.atol = float(overrides.get("atol", apg.get("atol", 0.0)))
    c.rtol = float(overrides.get("rtol", apg.get("rtol", 0.0)))
    c.beta_init = float(apg.get("beta_init", 0.25))
    ms = apg.get("moment_scale", None)
    c.moment_scale = 0.0 if ms is None else float(ms)
    if ms is not None and not 0.0 < float(ms) <= 1.0:
        raise ConfigError("apg_mpc.moment_scale must be null or in (0, 1] (launch/iris_sitl_traj_mpc.yaml:63-66)")
    if ms is None and c.beta_init != 0.25:
        # the momentum rule is beta_k = k / (k + 3) (launch/iris_sitl_traj_mpc.yaml:63-66), whose first value is 1/4
        msg = (f"apg_mpc.beta_init = {c.beta_init} is not applied while moment_scale is null: the classical momentum is "
               "beta_k = k / (k + 3) (beta_1 = 0.25)")
        if strict:
            raise ConfigError(msg)
        warnings.warn(msg)

    unsupported = [k for k in _UNSUPPORTED_COST_KEYS if k in cp] + [k for k in _UNSUPPORTED_TOP_KEYS if k in cfg]
    if unsupported:
        msg = f"config keys parsed but not applied by this solver: {unsupported} (see DESIGN.md, out of scope)"
        if strict:
            raise ConfigError(msg)
        warnings.warn(msg)
    return c


def config_to_dict(c: _abi.Config) -> dict:
    out = {}
    for name, _ in c._fields_:
        v = getattr(c, name)
        out[name] = list(v) if hasattr(v, "__len__") else v
    return out
