"""Independent float64 PyTorch statement of the MPC objective J(u) (SURVEY.md section 8a [SPEC]),
used to cross-check the oracle's hand-written adjoint with autograd.  Internal frame (NED/FRD)."""
import numpy as np
import torch


def unpack_model(model):
    t = lambda a: torch.tensor(np.asarray(a), dtype=torch.float64)
    w = {k: t(v) for k, v in model.weights.items()}
    f32 = lambda v: float(np.float32(v))   # the model blob stores float32 scalars
    return dict(w=w, m=f32(model.mass), g=f32(model.gravity), kT=f32(model.k_thrust), J=t(model.inertia), mix=t(model.mixer),
                sig0=t(model.sigma_prior))


def mlp(w, net, z):
    h = torch.tanh(w[f"{net}_W1"] @ z + w[f"{net}_b1"])
    h = torch.tanh(w[f"{net}_W2"] @ h + w[f"{net}_b2"])
    return w[f"{net}_W3"] @ h + w[f"{net}_b3"]


def rot(q):
    w, x, y, z = q
    return torch.stack([
        torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)]),
        torch.stack([2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)]),
        torch.stack([2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)])])


def qmul(a, b):
    return torch.stack([a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3],
                        a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
                        a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1],
                        a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0]])


def cost(cfg, M, x0, u, uprev, xref, xi):
    """cfg: _abi.Config; x0[13]; u[H,nu] (torch, requires grad); uprev[nu]; xref[H+1,13]; xi[P,H,6]."""
    H, P = cfg.horizon, cfg.num_particles
    T = lambda a: torch.tensor(list(a), dtype=torch.float64)
    perr, verr, qerr, werr, uref = T(cfg.perr), T(cfg.verr), T(cfg.qerr), T(cfg.werr), T(cfg.uref)[: cfg.nu]
    Jtot = 0.0
    for p in range(P):
        x = torch.tensor(x0, dtype=torch.float64)
        J, disc = 0.0, 1.0
        for t in range(H):
            dt = float(cfg.dt[t])
            pos, v, q, om = x[0:3], x[3:6], x[6:10], x[10:13]
            R = rot(q)
            z = torch.cat([R.T @ v, om, u[t]])
            r = mlp(M["w"], "drift", z)
            sig = M["sig0"] * torch.nn.functional.softplus(mlp(M["w"], "diff", z))
            Tm = M["kT"] * u[t] ** 2
            Fb = torch.stack([M["m"] * r[0], M["m"] * r[1], M["m"] * r[2] - Tm.sum()])
            Mb = M["mix"] @ Tm + M["J"] * r[3:6]
            acc = torch.tensor([0, 0, M["g"]], dtype=torch.float64) + R @ Fb / M["m"]
            qd = 0.5 * qmul(q, torch.cat([torch.zeros(1, dtype=torch.float64), om]))
            wd = (Mb - torch.linalg.cross(om, M["J"] * om)) / M["J"]
            n = torch.tensor(xi[p, t], dtype=torch.float64) * np.sqrt(dt)
            qn = q + qd * dt
            xn = torch.cat([pos + v * dt, v + acc * dt + sig[0:3] * n[0:3], qn / qn.norm(), om + wd * dt + sig[3:6] * n[3:6]])
            xr = torch.tensor(xref[t + 1], dtype=torch.float64)
            qr = xr[6:10]
            e = qmul(torch.stack([qr[0], -qr[1], -qr[2], -qr[3]]), xn[6:10])[1:]
            up = torch.tensor(uprev, dtype=torch.float64) if t == 0 else u[t - 1]
            l = (perr * (xn[0:3] - xr[0:3]) ** 2).sum() + (verr * (xn[3:6] - xr[3:6]) ** 2).sum() + (qerr * e ** 2).sum() \
                + (werr * (xn[10:13] - xr[10:13]) ** 2).sum() + cfg.uerr * ((u[t] - uref) ** 2).sum() \
                + cfg.u_slew_coeff * ((u[t] - up) ** 2).sum() + cfg.res_mult * (sig ** 2).sum()
            if cfg.u_slew_constr_coeff != 0.0:   # soft rate constraint: squared violation of [lo, hi]
                lo = torch.tensor(list(cfg.u_slew_lo[: cfg.nu]), dtype=torch.float64)
                hi = torch.tensor(list(cfg.u_slew_hi[: cfg.nu]), dtype=torch.float64)
                d = u[t] - up
                l = l + cfg.u_slew_constr_coeff * ((torch.relu(d - hi) - torch.relu(lo - d)) ** 2).sum()
            J = J + disc * l
            disc *= cfg.discount
            x = xn
        Jtot = Jtot + J
    return Jtot / P
