"""Independent cross-checks of the two parts of the oracle that round 1 only checked against themselves:

* R9, the accelerated projected-gradient loop: a SECOND float64 implementation, written here from the YAML
  comments of the reference (momentum beta_k = k/(k+3): launch/iris_sitl_traj_mpc.yaml:63-66; Armijo constants
  coef / decrease_factor / increase_factor / reset_option / maxls: :75-85; atol / rtol: :71-73; the input box: :8-11)
  and SURVEY.md section 8(a) steps 1-6, on top of the independent PyTorch statement of J(u) (tests/torch_ref.py)
  with autograd for the gradient.  It shares no code with oracle/sdempc_oracle_impl.h: different language,
  vectorised NumPy state, autograd instead of the hand-written adjoint, plain sums instead of SPEC-ARITH
  reductions.  Its per-iteration decision trace (f_y, J_trial, step size, trial count, accept, J_x, ||g||^2, k)
  must equal the float64 oracle's.
* the noise shaping: an independent NumPy Philox4x32-10 (vectorised uint64 arithmetic, checked against the
  Random123 known answers) and an independent Box-Muller in float64 libm on the same words.
"""
import numpy as np
import pytest

from conftest import make_setup
from oracle import oracle as O
from sde4mbrl_px4_b200 import model_io, synthetic


# ------------------------------------------------------------------------------------------------
# independent APG (SURVEY 8a [SPEC] "APG", steps 1-6)
# ------------------------------------------------------------------------------------------------
def independent_apg(cfg, J_and_grad, J_only, plan, stepsize, shift=True):
    """plan[H,nu] float64 in; returns (u*, trace[it,8], telemetry dict).  J_and_grad(u) -> (J, g), J_only(u) -> J."""
    H, nu = cfg.horizon, cfg.nu
    lo = np.array(cfg.u_lo[:nu], np.float64)[None, :]
    hi = np.array(cfg.u_hi[:nu], np.float64)[None, :]
    proj = lambda u: np.minimum(np.maximum(u, lo), hi)
    coef, dec_f, inc_f = np.float64(cfg.coef), np.float64(cfg.decrease_factor), np.float64(cfg.increase_factor)
    max_s, atol, rtol = np.float64(cfg.max_stepsize), np.float64(cfg.atol), np.float64(cfg.rtol)
    if shift:                                   # warm start: drop the applied control, repeat the last row
        plan = np.concatenate([plan[1:], plan[-1:]], axis=0)
    xk = proj(plan.copy())
    yk = xk.copy()
    s = np.float64(stepsize) if stepsize > 0 else np.float64(cfg.init_stepsize)
    k, no_improve, Jx = 1, 0, None
    rows, sum_ls, sum_s, init_cost = [], 0.0, 0.0, None
    for it in range(1, cfg.max_iter + 1):
        fy, g = J_and_grad(yk)                                          # step 1
        if it == 1:
            Jx = init_cost = fy
        if cfg.reset_option == 1:                                       # step 2
            s = min(s * inc_f, max_s)
        ok, n_ls, Jp, xp = False, 0, None, None
        for j in range(cfg.maxls + 1):                                  # step 3: projected Armijo
            xp = proj(yk - s * g)
            Jp = J_only(xp)
            n_ls = j + 1
            if Jp <= fy + coef * float(np.sum(g * (xp - yk))):
                ok = True
                break
            if j < cfg.maxls:
                s = s * dec_f
        sum_ls += n_ls
        sum_s += s
        accept = ok and Jp <= Jx                                        # step 4: monotone safeguard
        converged = False
        if accept:
            beta = k / (k + 3.0)
            yk = proj(xp + beta * (xp - xk))
            xk = xp
            Jprev, Jx = Jx, Jp
            k += 1
            no_improve = 0
            converged = abs(Jprev - Jx) <= atol + rtol * abs(Jprev) or Jx <= atol
        else:
            yk = xk.copy()
            k = 1
            no_improve += 1
        rows.append([fy, Jp, s, n_ls, float(accept), Jx, float(np.sum(g * g)), k])
        if no_improve >= cfg.max_no_improvement_iter or converged:     # step 6
            break
    n_it = len(rows)
    return xk, np.array(rows), dict(avg_linesearch=sum_ls / n_it, stepsize=s, num_steps=n_it, avg_stepsize=sum_s / n_it,
                                    init_cost=init_cost, opt_cost=Jx)


def enu2ned64(x):
    """float64 ENU/FLU -> NED/FRD ([SPEC] "State / frames"), written out with quaternion products:
    q_ned = q_r (x) q_enu (x) q_b with q_r = (0, s, s, 0), q_b = (0, 1, 0, 0), hemisphere qw >= 0."""
    x = np.asarray(x, np.float64)

    def qmul(a, b):
        return np.stack([a[..., 0] * b[..., 0] - a[..., 1] * b[..., 1] - a[..., 2] * b[..., 2] - a[..., 3] * b[..., 3],
                         a[..., 0] * b[..., 1] + a[..., 1] * b[..., 0] + a[..., 2] * b[..., 3] - a[..., 3] * b[..., 2],
                         a[..., 0] * b[..., 2] - a[..., 1] * b[..., 3] + a[..., 2] * b[..., 0] + a[..., 3] * b[..., 1],
                         a[..., 0] * b[..., 3] + a[..., 1] * b[..., 2] - a[..., 2] * b[..., 1] + a[..., 3] * b[..., 0]], axis=-1)

    s = np.sqrt(0.5)
    q = qmul(qmul(np.broadcast_to(np.array([0.0, s, s, 0.0]), x[..., 6:10].shape), x[..., 6:10]),
             np.broadcast_to(np.array([0.0, 1.0, 0.0, 0.0]), x[..., 6:10].shape))
    q = q * np.where(q[..., :1] < 0, -1.0, 1.0)
    return np.concatenate([x[..., [1, 0]], -x[..., 2:3], x[..., [4, 3]], -x[..., 5:6], q, x[..., 10:11], -x[..., 11:13]], axis=-1)


def _torch_objective(cfg, model, x0_int, uprev, xref_int, xi):
    import torch

    import torch_ref

    M = torch_ref.unpack_model(model)

    def J_and_grad(u):
        ut = torch.tensor(u, dtype=torch.float64, requires_grad=True)
        J = torch_ref.cost(cfg, M, x0_int, ut, uprev, xref_int, xi)
        J.backward()
        return float(J.detach()), ut.grad.numpy().copy()

    def J_only(u):
        with torch.no_grad():
            return float(torch_ref.cost(cfg, M, x0_int, torch.tensor(u, dtype=torch.float64), uprev, xref_int, xi))

    return J_and_grad, J_only


@pytest.mark.parametrize("vehicle,mode,P,iters,seed", [("iris", "traj", 1, 30, 3), ("iris", "pos", 1, 20, 4), ("hexa", "traj", 2, 12, 5)])
def test_oracle_apg_matches_independent_apg(vehicle, mode, P, iters, seed):
    """Decision trace of the float64 oracle == an independently written APG loop over torch autograd (R9)."""
    cfg, _, _ = make_setup(vehicle, mode, enu=True, num_particles=P, max_iter=iters, rtol=0.0, atol=0.0)
    model = model_io.synthetic_model(vehicle, seed=seed, weight_scale=0.4, bias_scale=0.2)
    blob = model.to_blob()
    o = O.Oracle(cfg, blob, "f64")
    pr = synthetic.batched_problems(1, cfg.horizon, np.array(cfg.dt[: cfg.horizon]), seed=seed)
    rng = np.random.default_rng(seed)
    plan = np.clip(np.array(cfg.uref[: cfg.nu]) + 0.05 * rng.standard_normal((1, cfg.horizon, cfg.nu)), 1e-4, 1)
    xi = rng.standard_normal((1, P, cfg.horizon, 6))
    _, i0 = o.reset(1)
    if mode == "pos":
        xdes = pr["xref_win"][:, 0].astype(np.float64)
        kw, xref_ext = dict(xdes=xdes), np.repeat(xdes[:, None, :], cfg.horizon + 1, axis=1)
    else:
        kw, xref_ext = dict(xref_win=pr["xref_win"]), pr["xref_win"].astype(np.float64)
    uo, _, info, trace = o.solve(pr["x"], plan, i0, xi=xi, want_trace=True, **kw)
    # the independent side works in the internal (NED/FRD) frame of torch_ref.cost
    x0_int = enu2ned64(pr["x"][0])
    xref_int = enu2ned64(xref_ext[0])
    Jg, Jo = _torch_objective(cfg, model, x0_int, plan[0, 0].copy(), xref_int, xi[0])
    ui, ti, tel = independent_apg(cfg, Jg, Jo, plan[0].copy(), float(i0[0, 1]))
    n_it = int(info[0, 2])
    assert n_it == tel["num_steps"] == ti.shape[0]
    to = trace[0, :n_it]
    # discrete decisions are identical
    assert np.array_equal(to[:, 3], ti[:, 3]), "line-search trial counts"
    assert np.array_equal(to[:, 4], ti[:, 4]), "accept / reject"
    assert np.array_equal(to[:, 7], ti[:, 7]), "momentum counter k"
    assert 0 < to[:, 4].sum() and (mode == "pos" or to[:, 4].sum() < n_it), "the trace exercises accepts (and rejects)"
    assert to[:, 3].max() > 1, "the trace exercises backtracking"
    # continuous quantities: two float64 implementations with different summation orders
    for col, name in ((0, "f_y"), (1, "J_trial"), (2, "step size"), (5, "J_x"), (6, "|g|^2")):
        err = np.abs(to[:, col] - ti[:, col]) / np.maximum(np.abs(ti[:, col]), 1e-300)
        assert err.max() <= 1e-9, (name, err.max())
    assert np.abs(uo[0] - ui).max() <= 1e-10
    assert abs(info[0, 0] - tel["avg_linesearch"]) <= 1e-12 and abs(info[0, 6] - tel["opt_cost"]) <= 1e-9 * abs(tel["opt_cost"])
    assert abs(info[0, 5] - tel["init_cost"]) <= 1e-9 * abs(tel["init_cost"]) and abs(info[0, 4] - tel["avg_stepsize"]) <= 1e-12


def test_oracle_apg_early_stop_and_no_improvement_match_independent_apg():
    """Stopping rules (yaml:60, 71-73): rtol stop and max_no_improvement_iter stop agree with the independent loop."""
    for over in (dict(rtol=3e-3, atol=0.0, max_iter=60), dict(rtol=0.0, atol=0.0, max_iter=60, max_no_improvement_iter=2)):
        cfg, _, _ = make_setup("iris", "traj", enu=False, **over)
        model = model_io.synthetic_model("iris", seed=9, weight_scale=0.3, bias_scale=0.1)
        o = O.Oracle(cfg, model.to_blob(), "f64")
        pr = synthetic.batched_problems(1, cfg.horizon, np.array(cfg.dt[: cfg.horizon]), seed=21)
        plan = np.full((1, cfg.horizon, cfg.nu), 0.71)
        xi = np.random.default_rng(2).standard_normal((1, 1, cfg.horizon, 6))
        _, i0 = o.reset(1)
        uo, _, info, trace = o.solve(pr["x"], plan, i0, xref_win=pr["xref_win"], xi=xi, want_trace=True)
        Jg, Jo = _torch_objective(cfg, model, pr["x"][0].astype(np.float64), plan[0, 0].copy(), pr["xref_win"][0].astype(np.float64), xi[0])
        ui, ti, tel = independent_apg(cfg, Jg, Jo, plan[0].copy(), float(i0[0, 1]))
        assert int(info[0, 2]) == tel["num_steps"] < cfg.max_iter, "stopped early, at the same iteration"
        assert np.array_equal(trace[0, : tel["num_steps"], 3:5], ti[:, 3:5]) and np.abs(uo[0] - ui).max() <= 1e-10


# ------------------------------------------------------------------------------------------------
# independent Philox + Box-Muller ([SPEC] "Noise")
# ------------------------------------------------------------------------------------------------
def np_philox4x32_10(ctr, key):
    """Vectorised Philox4x32-10 (Salmon et al., SC'11): ctr[...,4], key[...,2] uint32 -> [...,4] uint32."""
    c = [np.asarray(ctr[..., i], np.uint64) for i in range(4)]
    k = [np.asarray(key[..., i], np.uint64) for i in range(2)]
    M0, M1, W0, W1, MASK = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint64(0x9E3779B9), np.uint64(0xBB67AE85), np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        c = [(p1 >> np.uint64(32)) ^ c[1] ^ k[0], p1 & MASK, (p0 >> np.uint64(32)) ^ c[3] ^ k[1], p0 & MASK]
        k = [(k[0] + W0) & MASK, (k[1] + W1) & MASK]
    return np.stack(c, axis=-1).astype(np.uint32)


def np_noise(seed, tick, P, H, sub0=0):
    """xi[P,H,6] from ([SPEC] "Noise"): counter (t, particle, tick_lo, tick_hi(30 bits) | sub << 30), key = seed;
    words -> float32 uniforms in (0, 1] -> Box-Muller (cos, sin) in float64; 2 blocks, 6 of 8 normals."""
    t, p = np.meshgrid(np.arange(H, dtype=np.uint32), np.arange(P, dtype=np.uint32))
    key = np.broadcast_to(np.array([seed & 0xFFFFFFFF, seed >> 32], np.uint32), (P, H, 2))
    out = []
    for sub in (sub0, sub0 + 1):
        ctr = np.stack([t, p, np.full_like(t, tick & 0xFFFFFFFF), np.full_like(t, ((tick >> 32) & 0x3FFFFFFF) | (sub << 30))], axis=-1)
        w = np_philox4x32_10(ctr, key)
        # [SPEC]: u = (float32(w >> 8) + 0.5f) * 2^-24 evaluated in float32.  For w >> 8 >= 2^23 the half-cell offset
        # needs a 25th significand bit and rounds to even, so the upper half of (0, 1] is on the integer grid (u = 1
        # is reachable, u = 0 is not: log(u) stays finite).  Found by this independent check; it is the arithmetic
        # both the oracle and the kernels implement, so it is part of the noise specification.
        u = ((w >> np.uint32(8)).astype(np.float32) + np.float32(0.5)) * np.float32(2.0 ** -24)
        assert u.dtype == np.float32 and u.min() > 0.0 and u.max() <= 1.0
        u = u.astype(np.float64)
        for a, b in ((0, 1), (2, 3)):
            r = np.sqrt(-2.0 * np.log(u[..., a]))
            out += [r * np.cos(2 * np.pi * u[..., b]), r * np.sin(2 * np.pi * u[..., b])]
    return np.stack(out[:6], axis=-1)


def test_numpy_philox_known_answers():
    kat = [([0, 0, 0, 0], [0, 0], [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]),
           ([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2, [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]),
           ([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0], [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1])]
    for ctr, key, out in kat:
        assert [int(v) for v in np_philox4x32_10(np.array(ctr, np.uint32), np.array(key, np.uint32))] == out


@pytest.mark.parametrize("seed,tick,sub0", [(10, 0, 0), (1000 + 77, 123456789012, 0), ((1 << 40) + 5, 3, 2)])
def test_oracle_noise_matches_independent_box_muller(seed, tick, sub0):
    """The oracle's noise (C Philox + deterministic log / sincos sequences in f32, libm in f64) equals an
    independent NumPy generator: f64 to rounding, f32 to the accuracy of the deterministic functions."""
    cfg, blob, _ = make_setup()
    ref = np_noise(seed, tick, 8, 20, sub0)
    a64 = O.Oracle(cfg, blob, "f64").noise(seed, tick, P=8, H=20, sub0=sub0)
    a32 = O.Oracle(cfg, blob, "f32").noise(seed, tick, P=8, H=20, sub0=sub0)
    assert np.abs(a64 - ref).max() <= 1e-12
    assert np.abs(a32 - ref).max() <= 4e-6
    assert np.abs(ref).max() < 6.0 and abs(ref.mean()) < 0.15
