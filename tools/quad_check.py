#!/usr/bin/env python3
"""Quad-kernel bring-up: parity against the oracle on a small batch, then launch time at the bench shape
for the group kernel and both quad variants.  python tools/quad_check.py [--vehicle iris|hexa32] [--skip-parity]"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sde4mbrl_px4_b200 import config, model_io, solver, synthetic  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=4096)
ap.add_argument("--iters", type=int, default=200)
ap.add_argument("--skip-parity", action="store_true")
ap.add_argument("--variants", default="group,quad4,quad8")
a = ap.parse_args()

cfgd = config.load_yaml(os.path.join(ROOT, "configs", "iris_traj.yaml"))
blob = model_io.synthetic_model("iris").to_blob()
KW = {"group": dict(group=True), "quad4": dict(quad=4), "quad8": dict(quad=8), "warp": dict(sequential_ls=True)}

if not a.skip_parity:
    from oracle import oracle as O
    for B, it in ((5, 3), (37, 25)):
        cfg0 = config.build_config(cfgd, convert_to_enu=True, max_iter=it)
        pr = synthetic.batched_problems(B, cfg0.horizon, np.array(cfg0.dt[: cfg0.horizon]), seed=11 + B)
        o = O.Oracle(cfg0, blob, "f32")
        u0 = np.zeros((B, cfg0.horizon, cfg0.nu), np.float32)
        info0 = np.zeros((B, 8), np.float32)
        uo, xeo, infoo, tro = o.solve(pr["x"], u0, info0, xref_win=pr["xref_win"], rng=pr["rng"], want_trace=True)
        for name in a.variants.split(","):
            cfg = config.build_config(cfgd, convert_to_enu=True, max_iter=it, **KW[name])
            s = solver.MPCSolver(cfg, blob, device=0)
            u, xe, info, tr = s.solve(pr["x"], u0, info0, xref_win=pr["xref_win"], rng=pr["rng"], want_trace=True)
            ok = np.array_equal(u, uo) and np.array_equal(xe, xeo) and np.array_equal(tr, tro, equal_nan=True) \
                and np.array_equal(info[:, :7], infoo[:, :7])
            print(f"parity B={B} it={it} {name}: {'OK' if ok else 'MISMATCH'} max|du|={np.abs(u - uo).max():.3e} "
                  f"max|dx|={np.abs(xe - xeo).max():.3e} kernel={s.kernel_info()}", flush=True)
            if not ok:
                bad = np.argwhere(~np.isclose(tr, tro, rtol=0, atol=0, equal_nan=True))
                print("  first trace mismatches (problem, iter, field):", bad[:8].tolist())
                print("  gpu :", tr[tuple(bad[0][:2])] if len(bad) else None)
                print("  ref :", tro[tuple(bad[0][:2])] if len(bad) else None)

B = a.batch
for name in a.variants.split(","):
    cfg = config.build_config(cfgd, convert_to_enu=True, max_iter=a.iters, rtol=0.0, atol=0.0, **KW[name])
    pr = synthetic.batched_problems(B, cfg.horizon, np.array(cfg.dt[: cfg.horizon]), seed=1)
    s = solver.MPCSolver(cfg, blob, device=0)
    u0 = np.zeros((B, cfg.horizon, cfg.nu), np.float32)
    info0 = np.zeros((B, 8), np.float32)
    s.stage(pr["x"], u0, info0, xref_win=pr["xref_win"], rng=pr["rng"])
    s.launch_timed(2)
    ms = s.launch_timed(5)
    u, xe, info = s.fetch()
    print(f"time B={B} it={a.iters} {name}: {np.median(ms):.2f} ms  ({B / np.median(ms) * 1e3:.0f} solves/s) "
          f"mean n_ls {info[:, 0].mean():.3f} kernel={s.kernel_info()}", flush=True)
