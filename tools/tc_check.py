#!/usr/bin/env python3
"""Tensor-core rollout (SDEMPC_F_TENSOR) against the FP32 path: error statistics on a small batch, then the launch
time of both at a large batch.  python tools/tc_check.py [--batch 65536] [--lib path]"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sde4mbrl_px4_b200 import config, model_io, solver, synthetic  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=65536)
ap.add_argument("--vehicle", default="iris")
ap.add_argument("--lib", default=None)
ap.add_argument("--particles", type=int, default=1)
a = ap.parse_args()
cfgd = config.load_yaml(os.path.join(ROOT, "configs", f"{a.vehicle}_traj.yaml"))
blob = model_io.synthetic_model(a.vehicle).to_blob()
cfg_f = config.build_config(cfgd, convert_to_enu=True, num_particles=a.particles)
cfg_t = config.build_config(cfgd, convert_to_enu=True, tensor=True, num_particles=a.particles)
sf, st = solver.MPCSolver(cfg_f, blob, lib_path=a.lib), solver.MPCSolver(cfg_t, blob, lib_path=a.lib)
H, nu = cfg_f.horizon, cfg_f.nu


def problem(B, seed):
    pr = synthetic.batched_problems(B, H, np.array(cfg_f.dt[:H]), seed=seed)
    rng = np.random.default_rng(seed)
    u = np.clip(np.array(cfg_f.uref[:nu]) + 0.05 * rng.standard_normal((B, H, nu)), 1e-4, 1).astype(np.float32)
    up = np.tile(np.array(cfg_f.uref[:nu], np.float32), (B, 1))
    return pr, u, up


pr, u, up = problem(300, 5)
Jf, _, xf = sf.rollout(pr["x"], u, up, xref_win=pr["xref_win"], rng=pr["rng"], want_grad=False)
Jt, _, xt = st.rollout(pr["x"], u, up, xref_win=pr["xref_win"], rng=pr["rng"], want_grad=False)
rel = np.abs(Jt - Jf) / np.abs(Jf)
print(f"particles {a.particles}; B=300: cost rel err max {rel.max():.3e} median {np.median(rel):.3e}; x_evol max abs err {np.abs(xt - xf).max():.3e} "
      f"(|x| max {np.abs(xf).max():.2f}); J[0:3] fp32 {Jf[:3]} tensor {Jt[:3]}", flush=True)

Jf, gf, _ = sf.rollout(pr["x"], u, up, xref_win=pr["xref_win"], rng=pr["rng"], want_grad=True)
Jt, gt, _ = st.rollout(pr["x"], u, up, xref_win=pr["xref_win"], rng=pr["rng"], want_grad=True)
gerr = np.abs(gt - gf).reshape(300, -1).max(axis=1) / np.abs(gf).reshape(300, -1).max(axis=1)
print(f"B=300 value_and_grad: cost rel err max {(np.abs(Jt - Jf) / np.abs(Jf)).max():.3e}; gradient error / max|g| per problem: "
      f"max {gerr.max():.3e} median {np.median(gerr):.3e}; g[0,0] fp32 {gf[0, 0]} tensor {gt[0, 0]}", flush=True)

B = a.batch
pr, u, up = problem(B, 7)
for name, s in (("fp32", sf), ("tensor", st)):
    ms = []
    for _ in range(4):
        s.rollout(pr["x"], u, up, xref_win=pr["xref_win"], rng=pr["rng"], want_grad=False)
        ms.append(s.last_launch_ms())
    ms = np.array(ms[1:])
    print(f"{name}: B={B} forward rollouts (H={H}) {np.median(ms):.3f} ms -> {B / np.median(ms) * 1e3 / 1e6:.2f} M rollouts/s "
          f"(grid {s.kernel_info()['ctas']})", flush=True)

Bg = min(B, 65536)
pr, u, up = problem(Bg, 9)
for name, s in (("fp32", sf), ("tensor", st)):
    ms = []
    for _ in range(3):
        s.rollout(pr["x"], u, up, xref_win=pr["xref_win"], rng=pr["rng"], want_grad=True)
        ms.append(s.last_launch_ms())
    ms = np.array(ms[1:])
    print(f"{name}: B={Bg} value_and_grad {np.median(ms):.3f} ms -> {Bg / np.median(ms) * 1e3 / 1e6:.2f} M/s", flush=True)
