#!/usr/bin/env python3
"""One tensor-core batched cost evaluation / value_and_grad (SDEMPC_F_TENSOR) for `ncu` captures (profiles/README.md,
r1g, r1h).  python tools/tc_profile.py [vehicle] [problems] [grad: 0|1] [particles]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sde4mbrl_px4_b200 import config, model_io, solver, synthetic  # noqa: E402

vehicle = sys.argv[1] if len(sys.argv) > 1 else "iris"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
grad = len(sys.argv) > 3 and sys.argv[3] == "1"
particles = int(sys.argv[4]) if len(sys.argv) > 4 else 1
cfgd = config.load_yaml(os.path.join(ROOT, "configs", f"{vehicle}_traj.yaml"))
cfg = config.build_config(cfgd, tensor=True, num_particles=particles)
s = solver.MPCSolver(cfg, model_io.synthetic_model(vehicle).to_blob())
H, nu = cfg.horizon, cfg.nu
pr = synthetic.batched_problems(B, H, np.array(cfg.dt[:H]), seed=7)
u = np.full((B, H, nu), float(cfg.uref[0]), np.float32)
up = u[:, 0].copy()
for _ in range(2):
    s.rollout(pr["x"], u, up, xref_win=pr["xref_win"], rng=pr["rng"], want_grad=grad)
print("ms", s.last_launch_ms(), s.kernel_info())
