#!/usr/bin/env python3
"""One batched MPC solve on staged inputs, for `ncu` captures (see profiles/README.md)."""
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
ap.add_argument("--vehicle", default="iris")
ap.add_argument("--particles", type=int, default=1)
ap.add_argument("--launches", type=int, default=1)
ap.add_argument("--lib", default=None, help="alternative libsdempc build (tools/ab_variant.sh)")
a = ap.parse_args()
cfgd = config.load_yaml(os.path.join(ROOT, "configs", f"{a.vehicle}_traj.yaml"))
cfg = config.build_config(cfgd, max_iter=a.iters, rtol=0.0, atol=0.0, num_particles=a.particles)
blob = model_io.synthetic_model(a.vehicle).to_blob()
pr = synthetic.batched_problems(a.batch, cfg.horizon, np.array(cfg.dt[: cfg.horizon]), seed=0)
s = solver.MPCSolver(cfg, blob, lib_path=a.lib)
u0, i0 = s.reset(a.batch)
s.stage(pr["x"], u0, i0, xref_win=pr["xref_win"], rng=pr["rng"])
ms = s.launch_timed(a.launches, flush_l2=False)
u, xe, info = s.fetch()
print("ms", ms, "mean n_ls", info[:, 0].mean(), s.kernel_info())
