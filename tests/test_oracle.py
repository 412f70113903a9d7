"""CPU tests of the oracle (the parity checker itself): the externally pinned pieces
(Philox known-answer vectors), its internal consistency (adjoint vs finite differences and
vs an independent PyTorch autograd statement, f32 vs f64), the SPEC'd solver behaviour and
the committed golden vectors."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from conftest import ROOT, make_setup, random_states
from oracle import oracle as O
from sde4mbrl_px4_b200 import synthetic, trajectory
from sde4mbrl_px4_b200.utils import enu2ned


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32 with 10 rounds
    kat = [
        ([0, 0, 0, 0], [0, 0], [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]),
        ([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2, [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]),
        ([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0],
         [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]),
    ]
    for ctr, key, out in kat:
        assert [int(v) for v in O.philox4x32_10(ctr, key)] == out


def test_noise_is_standard_normal_and_f32_matches_f64():
    cfg, blob, _ = make_setup()
    o32, o64 = O.Oracle(cfg, blob, "f32"), O.Oracle(cfg, blob, "f64")
    xi = np.concatenate([o32.noise(7, k, P=64, H=20).reshape(-1) for k in range(16)])
    assert abs(xi.mean()) < 0.02 and abs(xi.std() - 1.0) < 0.02
    assert abs(np.mean(xi ** 3)) < 0.05 and abs(np.mean(xi ** 4) - 3.0) < 0.15
    a, b = o32.noise(123, 5, P=8, H=20), o64.noise(123, 5, P=8, H=20)
    assert np.abs(a - b).max() < 5e-6
    # sub-streams are distinct
    assert not np.array_equal(o32.noise(1, 0, sub0=0), o32.noise(1, 0, sub0=2))


@pytest.mark.parametrize("which,lo,hi,tol", [(0, -12, 12, 3e-7), (1, -30, 30, 5e-7), (2, -30, 30, 3e-7), (3, 0.7, 1.4, 3e-7)])
def test_det_math_accuracy(which, lo, hi, tol):
    """The deterministic float32 elementary functions agree with libm float64."""
    cfg, blob, _ = make_setup()
    o32, o64 = O.Oracle(cfg, blob, "f32"), O.Oracle(cfg, blob, "f64")
    x = np.linspace(lo, hi, 20001)
    y32 = np.zeros_like(x, dtype=np.float32)
    y64 = np.zeros_like(x)
    x32 = x.astype(np.float32)
    o32._fn("elementary", None)(which, x32.ctypes.data_as(C.POINTER(C.c_float)), y32.ctypes.data_as(C.POINTER(C.c_float)), len(x))
    x64 = x32.astype(np.float64)
    o64._fn("elementary", None)(which, x64.ctypes.data_as(C.POINTER(C.c_double)), y64.ctypes.data_as(C.POINTER(C.c_double)), len(x))
    err = np.abs(y32 - y64) / np.maximum(1.0, np.abs(y64))
    assert err.max() < tol, err.max()


def test_enu2ned_matches_python_and_is_involution():
    cfg, blob, _ = make_setup()
    o = O.Oracle(cfg, blob, "f32")
    x = random_states(64, 3)
    y = o.enu2ned(x)
    assert np.array_equal(y, enu2ned(x, np))
    z = o.enu2ned(y)
    sgn = np.sign(np.sum(z[:, 6:10] * x[:, 6:10], axis=1, keepdims=True))
    z[:, 6:10] *= sgn
    assert np.abs(z - x).max() < 1e-6
    assert np.all(y[:, 6] >= 0)
    # identity attitude in ENU/FLU (nose east) is a +90 deg yaw in NED/FRD
    e = o.enu2ned(np.array([[0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0]], np.float32))[0]
    assert np.allclose(e[6:10], [np.sqrt(0.5), 0, 0, np.sqrt(0.5)], atol=1e-7)


def _problem(cfg, B, seed):
    pr = synthetic.batched_problems(B, cfg.horizon, np.array(cfg.dt[: cfg.horizon]), seed=seed)
    rng = np.random.default_rng(seed)
    u = np.clip(np.array(cfg.uref[: cfg.nu]) + 0.05 * rng.standard_normal((B, cfg.horizon, cfg.nu)), 1e-4, 1).astype(np.float32)
    up = np.tile(np.array(cfg.uref[: cfg.nu], np.float32), (B, 1))
    return pr, u, up


@pytest.mark.parametrize("vehicle,mode,P", [("iris", "traj", 1), ("iris", "traj", 4), ("hexa", "traj", 2), ("iris", "pos", 1)])
def test_adjoint_vs_finite_differences(vehicle, mode, P):
    cfg, blob, _ = make_setup(vehicle, mode, num_particles=P)
    o = O.Oracle(cfg, blob, "f64")
    pr, u, up = _problem(cfg, 1, 11)
    u = u.astype(np.float64)
    J, g, _ = o.rollout(pr["x"], u, up, xref_win=pr["xref_win"], rng=pr["rng"])
    eps = 1e-6
    rng = np.random.default_rng(0)
    for _ in range(12):
        t, i = rng.integers(cfg.horizon), rng.integers(cfg.nu)
        d = np.zeros_like(u)
        d[0, t, i] = eps
        jp = o.rollout(pr["x"], u + d, up, xref_win=pr["xref_win"], rng=pr["rng"], want_grad=False)[0][0]
        jm = o.rollout(pr["x"], u - d, up, xref_win=pr["xref_win"], rng=pr["rng"], want_grad=False)[0][0]
        fd = (jp - jm) / (2 * eps)
        assert abs(fd - g[0, t, i]) <= 1e-6 * max(1.0, np.abs(g).max())


@pytest.mark.parametrize("vehicle,mode,P", [("iris", "traj", 2), ("hexa", "traj", 1), ("iris", "pos", 1)])
def test_cost_and_adjoint_vs_torch_autograd(vehicle, mode, P):
    """An independently written PyTorch float64 statement of J(u) gives the same cost and gradient."""
    import torch

    import torch_ref

    cfg, _, _ = make_setup(vehicle, mode, enu=False, num_particles=P)
    from sde4mbrl_px4_b200 import model_io
    model = model_io.synthetic_model(vehicle, weight_scale=0.3, bias_scale=0.2)   # non-zero biases on every layer
    blob = model.to_blob()
    o = O.Oracle(cfg, blob, "f64")
    pr, u, up = _problem(cfg, 1, 5)
    xi = np.random.default_rng(1).standard_normal((1, P, cfg.horizon, 6))
    J, g, _ = o.rollout(pr["x"], u.astype(np.float64), up, xref_win=pr["xref_win"], xi=xi)
    ut = torch.tensor(u[0].astype(np.float64), requires_grad=True)
    Jt = torch_ref.cost(cfg, torch_ref.unpack_model(model), pr["x"][0].astype(np.float64), ut, up[0], pr["xref_win"][0], xi[0])
    Jt.backward()
    assert abs(float(Jt) - J[0]) <= 1e-9 * abs(J[0])
    assert np.abs(ut.grad.numpy() - g[0]).max() <= 1e-8 * np.abs(g).max()


def test_soft_slew_constraint_is_active_and_one_sided():
    """iris_pos.yaml carries the reference's u_slew_constr (iris_sitl_posctrl_mpc.yaml:40-41): the penalty is
    zero inside [lo, hi], quadratic in the violation outside, and switched off by a zero coefficient."""
    cfg, blob, _ = make_setup("iris", "pos")
    assert cfg.u_slew_constr_coeff == 10.0 and abs(cfg.u_slew_hi[0] - 0.07) < 1e-7 and cfg.u_slew_lo[3] == -10.0
    o = O.Oracle(cfg, blob, "f64")
    cfg0, _, _ = make_setup("iris", "pos")
    cfg0.u_slew_constr_coeff = 0.0
    o0 = O.Oracle(cfg0, blob, "f64")
    pr, _, up = _problem(cfg, 1, 3)
    H, nu = cfg.horizon, cfg.nu
    base = np.tile(np.array(cfg.uref[:nu], np.float64), (1, H, 1))
    kw = dict(xref_win=pr["xref_win"], rng=pr["rng"], want_grad=False)
    # constant plan: no rate at all -> no penalty
    assert o.rollout(pr["x"], base, up, **kw)[0][0] == o0.rollout(pr["x"], base, up, **kw)[0][0]
    # a step of +0.1 on motor 0 at t = 5 violates hi = 0.07 by 0.03 once; the step back down (-0.1 at t = 6) is inside lo
    u = base.copy()
    u[0, 5, 0] += 0.1
    dJ = o.rollout(pr["x"], u, up, **kw)[0][0] - o0.rollout(pr["x"], u, up, **kw)[0][0]
    assert abs(dJ - 10.0 * 0.03 ** 2) < 1e-8   # hi is stored as float32
    # the same step on motor 1 (hi = 0.32) stays inside the band
    u = base.copy()
    u[0, 5, 1] += 0.1
    assert o.rollout(pr["x"], u, up, **kw)[0][0] == o0.rollout(pr["x"], u, up, **kw)[0][0]


def test_f32_oracle_close_to_f64():
    cfg, blob, _ = make_setup("iris", "traj", num_particles=2)
    o32, o64 = O.Oracle(cfg, blob, "f32"), O.Oracle(cfg, blob, "f64")
    pr, u, up = _problem(cfg, 8, 2)
    J32, g32, x32 = o32.rollout(pr["x"], u, up, xref_win=pr["xref_win"], rng=pr["rng"])
    J64, g64, x64 = o64.rollout(pr["x"], u, up, xref_win=pr["xref_win"], rng=pr["rng"])
    assert np.abs(J32 - J64).max() / np.abs(J64).max() < 2e-6
    assert np.abs(g32 - g64).max() / np.abs(g64).max() < 1e-5      # single gradient within 1e-5 rel
    assert np.abs(x32 - x64).max() < 1e-5


def test_hover_equilibrium_and_quaternion_norm():
    """With the residual networks zeroed and no noise, hover at uref is an equilibrium of the
    plant (k_T = m g / (nu uref^2)); the rollout keeps |q| = 1."""
    cfg, _, model = make_setup("iris", "pos", enu=False)
    for k in model.weights:
        model.weights[k][...] = 0
    o = O.Oracle(cfg, model.to_blob(), "f64")
    x = np.array([[1, 2, -3, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0]], np.float64)
    u = np.tile(np.array(cfg.uref[:4]), (1, cfg.horizon, 1))
    xi = np.zeros((1, 1, cfg.horizon, 6))
    J, _, xe = o.rollout(x, u, u[:, 0], xdes=x, xi=xi)
    assert np.abs(xe[0] - x[0]).max() < 1e-6     # uref is stored as float32, hence the 1e-7-level residual
    cfg2, blob2, _ = make_setup("iris", "traj", enu=False)
    o2 = O.Oracle(cfg2, blob2, "f32")
    pr, u, up = _problem(cfg2, 4, 9)
    xs = random_states(4, 9)
    _, _, xe = o2.rollout(xs, u, up, xdes=xs, rng=pr["rng"])
    assert np.abs(np.linalg.norm(xe[:, :, 6:10], axis=2) - 1).max() < 3e-7


def test_trajectory_interpolation_and_window():
    cfg, blob, _ = make_setup("iris", "traj", enu=True)
    o = O.Oracle(cfg, blob, "f32")
    tab = trajectory.csv_rows_to_table(trajectory.lemniscate(2.0, 8.0, 0.3, duration=10.0))
    o.set_trajectory(tab)
    t = np.array([-1.0, 0.0, 0.004, 1.2345, 9.999, 10.0, 25.0], np.float32)
    got = o.state_from_traj(t)
    ref = trajectory.interp_table(tab, t)
    assert np.abs(got - ref).max() < 1e-6
    assert np.array_equal(got[0], got[1]) and np.array_equal(got[-1], got[-2])     # clamped at both ends
    # internal table = enu2ned of the external rows
    ti = o.traj_internal()
    assert np.array_equal(ti[:, 0], tab[:, 0])
    assert np.abs(ti[:, 1:] - enu2ned(tab[:, 1:], np)).max() < 1e-6
    # trajectory mode == explicit window built from state_from_traj on the same time grid
    x = random_states(2, 4)
    u0, i0 = o.reset(2)
    ct = np.array([0.5, 3.25], np.float32)
    grid = np.concatenate([[0.0], np.cumsum(np.array(cfg.dt[: cfg.horizon], np.float32), dtype=np.float32)]).astype(np.float32)
    win = np.stack([o.state_from_traj((c + grid).astype(np.float32)) for c in ct])
    rng = np.array([[5, 0], [6, 0]], np.uint64)
    Ja, ga, _ = o.rollout(x, u0, u0[:, 0], curr_t=ct, rng=rng)
    Jb, gb, _ = o.rollout(x, u0, u0[:, 0], xref_win=win, rng=rng)
    assert np.abs(Ja - Jb).max() / np.abs(Jb).max() < 1e-5


def test_apg_spec_behaviour():
    """Monotone safeguard (J_x never increases), projection onto the input box, telemetry
    consistency, iteration budget, warm-start shift and rng-independence of a P=1 solve's structure."""
    cfg, blob, _ = make_setup("iris", "traj", max_iter=60, rtol=0.0, atol=0.0)
    o = O.Oracle(cfg, blob, "f32")
    pr, _, _ = _problem(cfg, 6, 21)
    u0, i0 = o.reset(6)
    assert np.allclose(u0, 0.71) and np.allclose(i0[:, 1], cfg.init_stepsize)
    u, xe, info, tr = o.solve(pr["x"], u0, i0, xref_win=pr["xref_win"], rng=pr["rng"], want_trace=True)
    assert np.all(info[:, 2] == 60)
    assert np.all(u >= 1e-4 - 1e-9) and np.all(u <= 1.0)
    Jx = tr[:, :, 5]
    assert np.all(np.diff(Jx, axis=1) <= 0)
    assert np.allclose(info[:, 6], Jx[:, -1]) and np.allclose(info[:, 5], tr[:, 0, 0])
    assert np.all(info[:, 6] < info[:, 5])
    assert np.allclose(info[:, 0], tr[:, :, 3].mean(axis=1))
    assert np.all((tr[:, :, 3] >= 1) & (tr[:, :, 3] <= cfg.maxls + 1))
    # the reported opt_cost is the cost of the returned plan (u_prev = first control of the incoming plan)
    J, _, xe2 = o.rollout(pr["x"], u, u0[:, 0], xref_win=pr["xref_win"], rng=pr["rng"], want_grad=False)
    assert np.array_equal(J, info[:, 6]) and np.array_equal(xe2, xe)
    assert np.abs(xe[:, 0] - pr["x"]).max() < 1e-6     # through ENU->NED->renormalise->ENU
    # warm start: the second tick starts from the shifted plan and a carried step size
    u2, _, info2, tr2 = o.solve(xe[:, 1], u, info, xref_win=pr["xref_win"], rng=pr["rng"], want_trace=True)
    shifted = np.concatenate([u[:, 1:], u[:, -1:]], axis=1)
    J0, _, _ = o.rollout(xe[:, 1], shifted, u[:, 0], xref_win=pr["xref_win"], rng=pr["rng"], want_grad=False)
    assert np.array_equal(J0, info2[:, 5])
    # early stopping with the YAML tolerances never exceeds the budget
    cfg3, blob3, _ = make_setup("iris", "pos")
    o3 = O.Oracle(cfg3, blob3, "f32")
    x = random_states(3, 1)
    u0, i0 = o3.reset(3)
    _, _, info3, _ = o3.solve(x, u0, i0, xdes=x, rng=pr["rng"][:3])
    assert np.all(info3[:, 2] <= cfg3.max_iter) and np.all(info3[:, 2] >= 1)


def test_closed_loop_stabilises_the_plant():
    """BASELINE config 5 in miniature: the spec'd controller tracks the lemniscate with the spec'd plant."""
    cfg, blob, _ = make_setup("iris", "traj", max_iter=30, rtol=0.0, atol=0.0)
    o = O.Oracle(cfg, blob, "f32")
    tab = trajectory.csv_rows_to_table(trajectory.lemniscate(1.5, 10.0, 0.0, duration=20.0))
    o.set_trajectory(tab)
    x0 = synthetic.initial_states(tab[0, 1:4], 3, seed=2)
    rng = np.array([[100 + r, 0] for r in range(3)], np.uint64)
    xh, uh, stats = o.closed_loop(x0, np.zeros(3, np.float32), rng, ticks=60)
    ref = trajectory.interp_table(tab, (np.arange(61) * 0.05).astype(np.float32))
    err = np.linalg.norm(xh[:, :, 0:3] - ref[None, :, 0:3], axis=2)
    assert np.all(err[:, -10:].max(axis=1) < 0.15), err[:, -10:].max(axis=1)
    assert np.all(err[:, -1] < err[:, 0])
    assert np.all(np.isfinite(stats))


GOLDEN = os.path.join(ROOT, "tests", "golden", "golden_v1.npz")


@pytest.mark.skipif(not os.path.exists(GOLDEN), reason="golden fixtures not generated")
def test_oracle_reproduces_golden_vectors():
    """The committed fixtures were written by tests/golden/make_golden.py from the oracle;
    a change in SPEC-ARITH shows up here first (bit-exact comparison)."""
    from golden import make_golden

    z = np.load(GOLDEN)
    meta = json.loads(str(z["meta"]))
    for case in meta["cases"]:
        out = make_golden.run_case(case, backend="oracle")
        for k, v in out.items():
            assert np.array_equal(v, z[f"{case['name']}/{k}"]), (case["name"], k)


def test_teacher_forced_mode_replays_a_decision_trace():
    """TEACHER-FORCED mode: forcing the oracle with its own decision trace reproduces the free
    solve bit for bit; forcing the float64 oracle with the float32 trace keeps the two on the same sequence of
    iterates, so early iterations agree to rounding, `own` reports the decisions it would have taken itself, and
    a deliberately wrong decision is followed, not corrected."""
    cfg, blob, _ = make_setup("iris", "traj", max_iter=40, rtol=0.0, atol=0.0)
    o32, o64 = O.Oracle(cfg, blob, "f32"), O.Oracle(cfg, blob, "f64")
    B = 12
    pr = synthetic.batched_problems(B, cfg.horizon, np.array(cfg.dt[: cfg.horizon]), seed=2)
    u0, i0 = o32.reset(B)
    kw = dict(xref_win=pr["xref_win"], rng=pr["rng"])
    a = o32.solve(pr["x"], u0, i0, want_trace=True, **kw)
    f = o32.solve_forced(pr["x"], u0, i0, a[3], a[2][:, 2], **kw)
    assert np.array_equal(f[0], a[0]) and np.array_equal(f[1], a[1]) and np.array_equal(f[3], a[3]) and np.array_equal(f[2][:, :7], a[2][:, :7])
    assert np.array_equal(f[4][:, :, 0], a[3][:, :, 3]) and np.array_equal(f[4][:, :, 1], a[3][:, :, 4]), "own decisions == forced decisions"
    acc = a[3][:, :, 4] > 0
    assert np.all(f[4][:, :, 2][acc] >= 0) and np.all(f[4][:, :, 3][acc] >= 0), "margins of accepted steps are non-negative"
    g = o64.solve_forced(pr["x"], u0, i0, a[3], a[2][:, 2], **kw)
    rel = np.abs(g[3][:, :10, [0, 5]] - a[3][:, :10, [0, 5]]) / np.abs(g[3][:, :10, [0, 5]])
    assert rel.max() <= 2e-5 and np.array_equal(g[3][:, :, 3:5], a[3][:, :, 3:5])
    # a wrong forced decision (reject the first accepted step of problem 0) is followed
    tr = a[3].copy()
    k0 = int(np.argmax(tr[0, :, 4] > 0))
    tr[0, k0, 4] = 0.0
    h = o32.solve_forced(pr["x"], u0, i0, tr, a[2][:, 2], **kw)
    assert h[3][0, k0, 4] == 0.0 and h[4][0, k0, 1] == 1.0 and h[3][0, k0, 7] == 1.0
    assert np.array_equal(h[0][1:], a[0][1:]) and not np.array_equal(h[0][0], a[0][0])
    # fewer forced iterations than max_iter: stops there
    short = o32.solve_forced(pr["x"], u0, i0, a[3], np.full(B, 7.0), **kw)
    assert np.all(short[2][:, 2] == 7) and np.array_equal(short[3][:, :7], a[3][:, :7])


def test_adaptive_momentum_and_unenforced_bounds_spec():
    """apg_mpc.moment_scale (yaml:63-66) and enforce_ubound: False (yaml:14), [SPEC] in include/sdempc.h: with mu in (0, 1]
    the momentum is beta_init / mu^(k-1) capped at 1 (mu = 1: constant beta_init); null keeps k / (k + 3).  The decision
    trace of the oracle is checked against the independent float64 APG of test_independent_apg with the same rule."""
    import os

    from conftest import ROOT
    from sde4mbrl_px4_b200 import config, model_io
    from test_independent_apg import _torch_objective, enu2ned64

    d = config.load_yaml(os.path.join(ROOT, "configs", "iris_traj.yaml"))
    d["apg_mpc"].update(moment_scale=0.6, beta_init=0.2, max_iter=25, rtol=0.0, atol=0.0)
    cfg = config.build_config(d, convert_to_enu=False, strict=True)
    assert abs(cfg.moment_scale - 0.6) < 1e-7 and abs(cfg.beta_init - 0.2) < 1e-7
    with pytest.raises(config.ConfigError):
        config.build_config(dict(d, apg_mpc=dict(d["apg_mpc"], moment_scale=1.5)))
    model = model_io.synthetic_model("iris", seed=2, weight_scale=0.3, bias_scale=0.1)
    o = O.Oracle(cfg, model.to_blob(), "f64")
    pr = synthetic.batched_problems(1, cfg.horizon, np.array(cfg.dt[: cfg.horizon]), seed=8)
    plan = np.full((1, cfg.horizon, cfg.nu), 0.71)
    xi = np.random.default_rng(4).standard_normal((1, 1, cfg.horizon, 6))
    _, i0 = o.reset(1)
    uo, _, info, tr = o.solve(pr["x"], plan, i0, xref_win=pr["xref_win"], xi=xi, want_trace=True)
    # momentum visible in the trace: after an accepted step with counter k -> k + 1, y = clip(x+ + beta_k (x+ - x_k))
    mu, b0 = np.float64(np.float32(0.6)), np.float64(np.float32(0.2))
    beta = lambda k: min(1.0, b0 / mu ** (k - 1)) if k < 60 else 1.0

    # independent loop with the same momentum rule
    import test_independent_apg as T
    Jg, Jo = _torch_objective(cfg, model, pr["x"][0].astype(np.float64), plan[0, 0].copy(), pr["xref_win"][0].astype(np.float64), xi[0])
    H, nu = cfg.horizon, cfg.nu
    lo, hi = np.array(cfg.u_lo[:nu], np.float64)[None], np.array(cfg.u_hi[:nu], np.float64)[None]
    proj = lambda u: np.minimum(np.maximum(u, lo), hi)
    p = np.concatenate([plan[0, 1:], plan[0, -1:]], axis=0)
    xk = proj(p.copy()); yk = xk.copy()
    s, k, Jx, rows = np.float64(i0[0, 1]), 1, None, []
    for it in range(1, cfg.max_iter + 1):
        fy, g = Jg(yk)
        Jx = fy if it == 1 else Jx
        s = min(s * np.float64(cfg.increase_factor), np.float64(cfg.max_stepsize))
        ok = False
        for j in range(cfg.maxls + 1):
            xp = proj(yk - s * g); Jp = Jo(xp)
            if Jp <= fy + np.float64(cfg.coef) * float(np.sum(g * (xp - yk))):
                ok = True; break
            if j < cfg.maxls:
                s = s * np.float64(cfg.decrease_factor)
        if ok and Jp <= Jx:
            yk = proj(xp + beta(k) * (xp - xk)); xk = xp; Jx = Jp; k += 1
        else:
            yk = xk.copy(); k = 1
        rows.append([fy, Jp, s, j + 1, float(ok and Jp <= Jx + 0), Jx, 0, k])
    ti = np.array(rows)
    n_it = int(info[0, 2])
    assert n_it == cfg.max_iter and np.array_equal(tr[0, :, 3], ti[:, 3]) and np.array_equal(tr[0, :, 7], ti[:, 7])
    assert np.abs(tr[0, :, 0] - ti[:, 0]).max() <= 1e-9 * np.abs(ti[:, 0]).max() and np.abs(uo[0] - xk).max() <= 1e-10
    assert tr[0, :, 7].max() >= 4, "several consecutive accepted steps: the momentum actually grew"
    # it differs from the classical rule
    d2 = dict(d, apg_mpc=dict(d["apg_mpc"], moment_scale=None, beta_init=0.25))
    o2 = O.Oracle(config.build_config(d2, convert_to_enu=False), model.to_blob(), "f64")
    u2 = o2.solve(pr["x"], plan, i0, xref_win=pr["xref_win"], xi=xi)[0]
    assert np.abs(u2 - uo).max() > 1e-6
    # enforce_ubound: False -> no projection: a plan outside the box stays outside when the gradient does not pull it in
    d3 = dict(d2, enforce_ubound=False)
    c3 = config.build_config(d3, convert_to_enu=False, max_iter=3)
    assert c3.u_lo[0] < -1e37 and c3.u_hi[0] > 1e37
    o3 = O.Oracle(c3, model.to_blob(), "f64")
    big = np.full((1, cfg.horizon, cfg.nu), 1.2)
    u3 = o3.solve(pr["x"], big, i0, xref_win=pr["xref_win"], xi=xi)[0]
    u3b = O.Oracle(config.build_config(d2, convert_to_enu=False, max_iter=3), model.to_blob(), "f64").solve(pr["x"], big, i0, xref_win=pr["xref_win"], xi=xi)[0]
    assert u3.max() > 1.0 + 1e-3 and u3b.max() <= 1.0
