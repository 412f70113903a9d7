#!/usr/bin/env python3
"""Launch time of the tensor-core rollout (cost evaluation and value_and_grad) against the number of rows, i.e.
against how many 128-row CTAs share an SM: the latency of a lone CTA is what bounds the tensor-core solve at small
batches, the full-occupancy figure is what bounds it at large ones.  python tools/tc_fill_sweep.py [vehicle]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sde4mbrl_px4_b200 import config, model_io, solver, synthetic  # noqa: E402

vehicle = sys.argv[1] if len(sys.argv) > 1 else "iris"
cfgd = config.load_yaml(os.path.join(ROOT, "configs", f"{vehicle}_traj.yaml"))
blob = model_io.synthetic_model(vehicle).to_blob()
cfg = config.build_config(cfgd, convert_to_enu=True, tensor=True)
s = solver.MPCSolver(cfg, blob)
H, nu = cfg.horizon, cfg.nu
for B in (128, 1024, 4096, 148 * 128, 2 * 148 * 128, 4 * 148 * 128, 65536 * 2):
    pr = synthetic.batched_problems(B, H, np.array(cfg.dt[:H]), seed=1)
    u = np.full((B, H, nu), float(cfg.uref[0]), np.float32)
    up = u[:, 0].copy()
    out = []
    for grad in (False, True):
        ms = []
        for _ in range(5):
            s.rollout(pr["x"], u, up, xref_win=pr["xref_win"], rng=pr["rng"], want_grad=grad)
            ms.append(s.last_launch_ms())
        out.append(float(np.median(ms[2:])))
    ctas = (B + 127) // 128
    print(f"{vehicle} rows {B:7d} CTAs {ctas:5d} ({ctas / 148:5.2f} per SM)  cost {out[0]:7.3f} ms = {out[0] / H * 1e3:6.2f} us/step   "
          f"value_and_grad {out[1]:7.3f} ms = {out[1] / H * 1e3:6.2f} us/step", flush=True)
