#!/usr/bin/env python3
"""Device time of one batched iris solve (200 iterations) vs batch size, for each kernel choice."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sde4mbrl_px4_b200 import config, model_io, solver, synthetic
vehicle = sys.argv[1] if len(sys.argv) > 1 else "iris"
width = int(sys.argv[2]) if len(sys.argv) > 2 else None
cfgd = config.load_yaml(os.path.join(ROOT, "configs", f"{vehicle}_traj.yaml"))
blob = model_io.synthetic_model(vehicle, width=width).to_blob()
for B in ((150, 296, 592, 1184, 1776, 2368, 3552, 4096, 8192) if vehicle == "iris" else (592, 1184, 2368, 4096)):
    row = []
    for mode in ("sequential_ls", "group", "speculative_ls"):
        if mode == "speculative_ls" and B > 592:
            row.append(float("nan")); continue
        cfg = config.build_config(cfgd, max_iter=200, rtol=0.0, atol=0.0, **{mode: True})
        pr = synthetic.batched_problems(B, cfg.horizon, np.array(cfg.dt[: cfg.horizon]), seed=0)
        s = solver.MPCSolver(cfg, blob)
        u0, i0 = s.reset(B)
        s.stage(pr["x"], u0, i0, xref_win=pr["xref_win"], rng=pr["rng"])
        ms = s.launch_timed(3, flush_l2=False)
        row.append(float(ms[1:].mean()))
    print(f"B {B:5d}  one-warp-per-problem {row[0]:8.2f} ms   group {row[1]:8.2f} ms   latency(8 warps/problem) {row[2]:8.2f} ms", flush=True)
