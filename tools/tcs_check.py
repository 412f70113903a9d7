#!/usr/bin/env python3
"""Tensor-core batched solve (SDEMPC_F_TENSOR through sdempc_solve_ex) against the oracle and the FP32 kernels:
teacher-forced comparison of the decision trace, free-running cost agreement, and launch times.
python tools/tcs_check.py [--vehicle iris] [--particles 1] [--iters 200] [--batches 4096,65536] [--scale 0.6]"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402  (checker only)
from sde4mbrl_px4_b200 import config, model_io, solver, synthetic  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--vehicle", default="iris")
ap.add_argument("--particles", type=int, default=1)
ap.add_argument("--iters", type=int, default=200)
ap.add_argument("--batches", default="4096,65536")
ap.add_argument("--scale", type=float, default=None, help="weight scale of the synthetic model (None: BASELINE default 0.1)")
ap.add_argument("--check", type=int, default=64)
ap.add_argument("--no-fp32", action="store_true")
a = ap.parse_args()

cfgd = config.load_yaml(os.path.join(ROOT, "configs", f"{a.vehicle}_traj.yaml"))
kw = dict(weight_scale=a.scale, bias_scale=0.2) if a.scale else {}
blob = model_io.synthetic_model(a.vehicle, **kw).to_blob()
ov = dict(convert_to_enu=True, num_particles=a.particles, max_iter=a.iters, rtol=0.0, atol=0.0)
cfg_f = config.build_config(cfgd, **ov)
cfg_t = config.build_config(cfgd, tensor=True, **ov)
H, nu = cfg_f.horizon, cfg_f.nu
st, sf = solver.MPCSolver(cfg_t, blob), solver.MPCSolver(cfg_f, blob)
o = O.Oracle(cfg_f, blob, "f32")

B = a.check
pr = synthetic.batched_problems(B, H, np.array(cfg_f.dt[:H]), seed=3)
u0, i0 = st.reset(B)
ut, xt, it_, trt = st.solve(pr["x"], u0, i0, xref_win=pr["xref_win"], rng=pr["rng"], want_trace=True)
print("tensor solve:", st.kernel_info(), "launch ms", st.last_launch_ms(), flush=True)
uo, xo, io, tro = o.solve(pr["x"], u0, i0, xref_win=pr["xref_win"], rng=pr["rng"], want_trace=True)
uf, xf, if_, trf, own = o.solve_forced(pr["x"], u0, i0, trt, it_[:, 2], xref_win=pr["xref_win"], rng=pr["rng"])
nit = it_[:, 2].astype(int)
print(f"iterations tensor {nit.min()}..{nit.max()}  oracle {int(io[:, 2].min())}..{int(io[:, 2].max())}; mean n_ls tensor {it_[:, 0].mean():.3f} oracle {io[:, 0].mean():.3f}")
# teacher-forced: same decisions, compare continuous quantities per iteration
m = np.arange(cfg_f.max_iter)[None, :] < nit[:, None]
for col, name in ((0, "f_y"), (5, "J_x"), (6, "|g|^2"), (2, "step")):
    d = np.abs(trt[:, :, col] - trf[:, :, col]) / np.maximum(np.abs(trf[:, :, col]), 1e-30)
    print(f"  teacher-forced {name:6s}: max rel {d[m].max():.3e}  median {np.median(d[m]):.3e}")
for kmax in (5, 10, 20, 50):
    mk = m & (np.arange(cfg_f.max_iter)[None, :] < kmax)
    d0 = np.abs(trt[:, :, 0] - trf[:, :, 0]) / np.maximum(np.abs(trf[:, :, 0]), 1e-30)
    d6 = np.abs(trt[:, :, 6] - trf[:, :, 6]) / np.maximum(np.abs(trf[:, :, 6]), 1e-30)
    print(f"  first {kmax:3d} iterations: f_y max rel {d0[mk].max():.3e} (99 % {np.quantile(d0[mk], 0.99):.3e}); |g|^2 max rel {d6[mk].max():.3e}; "
          f"flipped decisions {int(((own[:, :, 0] != trt[:, :, 3]) & mk).sum())} + {int(((own[:, :, 1] != trt[:, :, 4]) & mk).sum())} of {int(mk.sum())}")
d0 = np.abs(trt[:, :, 0] - trf[:, :, 0]) / np.maximum(np.abs(trf[:, :, 0]), 1e-30)
print("  f_y rel diff quantiles over all (problem, iteration): " + ", ".join(f"{q:.2f}: {np.quantile(d0[m], q):.2e}" for q in (0.5, 0.9, 0.99, 0.999)))
acc = m & (trt[:, :, 4] > 0)
d = np.abs(trt[:, :, 1] - trf[:, :, 1]) / np.maximum(np.abs(trf[:, :, 1]), 1e-30)
print(f"  teacher-forced J_trial on accepted steps: max rel {d[acc].max():.3e} median {np.median(d[acc]):.3e}")
flips_ls = (own[:, :, 0] != trt[:, :, 3]) & m
flips_acc = (own[:, :, 1] != trt[:, :, 4]) & m
print(f"  decisions the oracle would have taken differently on the same iterates: trial count {flips_ls.sum()} / {m.sum()}, accept {flips_acc.sum()} / {m.sum()}")
print(f"  teacher-forced u*: max abs diff {np.abs(ut - uf).max():.3e}; x_evol max abs diff {np.abs(xt - xf).max():.3e}")
rc = np.abs(it_[:, 6] - io[:, 6]) / np.abs(io[:, 6])
print(f"free-running: opt_cost rel diff vs oracle: max {rc.max():.3e} median {np.median(rc):.3e}; u* max abs diff {np.abs(ut - uo).max():.3e} median {np.median(np.abs(ut - uo).reshape(B, -1).max(axis=1)):.3e}")
print(f"  init_cost rel diff max {(np.abs(it_[:, 5] - io[:, 5]) / np.abs(io[:, 5])).max():.3e}; opt/init tensor {np.median(it_[:, 6] / it_[:, 5]):.4f} oracle {np.median(io[:, 6] / io[:, 5]):.4f}", flush=True)

for Bs in [int(v) for v in a.batches.split(",") if v]:
    pr = synthetic.batched_problems(Bs, H, np.array(cfg_f.dt[:H]), seed=0)
    for name, s in (("tensor", st),) + (() if a.no_fp32 else (("fp32", sf),)):
        u0, i0 = s.reset(Bs)
        s.stage(pr["x"], u0, i0, xref_win=pr["xref_win"], rng=pr["rng"])
        ms = s.launch_timed(3, flush_l2=True)
        _, _, info = s.fetch()
        ki = s.kernel_info()
        print(f"{name:6s} B={Bs:6d} P={a.particles}: {ms[1:].mean():9.3f} ms  -> {Bs / ms[1:].mean() * 1e3:10.0f} solves/s  (grid {ki['ctas']}, {ki['problems_per_cta']} problems/CTA, "
              f"mean n_ls {info[:, 0].mean():.3f}, median opt_cost {np.median(info[:, 6]):.3f})", flush=True)
