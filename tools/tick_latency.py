#!/usr/bin/env python3
"""Single-tick latency of one MPC solve (B = 1) for a vehicle / particle count, warm-started along a lemniscate:
    python tools/tick_latency.py [--vehicle hexa] [--particles 8] [--ticks 100] [--iters 200]
Prints p50 / p99 of the end-to-end call (host buffers) and the kernel that ran."""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sde4mbrl_px4_b200 import config, model_io, solver, trajectory  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--vehicle", default="hexa")
ap.add_argument("--particles", type=int, default=8)
ap.add_argument("--ticks", type=int, default=100)
ap.add_argument("--iters", type=int, default=200)
ap.add_argument("--batch", type=int, default=1)
a = ap.parse_args()
cfgd = config.load_yaml(os.path.join(ROOT, "configs", f"{a.vehicle}_traj.yaml"))
cfg = config.build_config(cfgd, max_iter=a.iters, rtol=0.0, atol=0.0, num_particles=a.particles)
s = solver.MPCSolver(cfg, model_io.synthetic_model(a.vehicle).to_blob())
tab = trajectory.csv_rows_to_table(trajectory.lemniscate(2.0, 8.0, 0.0, duration=60.0))
s.set_trajectory(tab)
B = a.batch
x = np.repeat(tab[0:1, 1:], B, 0).astype(np.float32)
up, ip = s.reset(B)
ts = []
for k in range(a.ticks + 5):
    t = 0.05 * k
    rng = np.array([[3, k]] * B, np.uint64)
    t0 = time.perf_counter()
    up, xe, ip, _ = s.solve(x, up, ip, curr_t=np.full(B, t, np.float32), rng=rng)
    ts.append((time.perf_counter() - t0) * 1e3)
    x = xe[:, 1].copy()
ts = np.array(ts[5:])
print(f"{a.vehicle} P={a.particles} B={B}: p50 {np.percentile(ts, 50):.3f} ms  p99 {np.percentile(ts, 99):.3f} ms  "
      f"mean iterations {ip[:, 2].mean():.1f}  kernel {s.kernel_info()}")
