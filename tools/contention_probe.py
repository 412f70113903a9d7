#!/usr/bin/env python3
"""How does per-warp solve time change with the number of co-resident warps on one SM?"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sde4mbrl_px4_b200 import config, model_io, solver, synthetic
cfgd = config.load_yaml(os.path.join(ROOT, "configs", "iris_traj.yaml"))
blob = model_io.synthetic_model("iris").to_blob()
for mode in ("sequential_ls", "speculative_ls"):
    cfg = config.build_config(cfgd, max_iter=200, rtol=0.0, atol=0.0, **{mode: True})
    for B in (1, 2, 3, 4, 5, 6, 8):
        pr = synthetic.batched_problems(B, cfg.horizon, np.array(cfg.dt[: cfg.horizon]), seed=0)
        s = solver.MPCSolver(cfg, blob)
        u0, i0 = s.reset(B)
        s.stage(pr["x"], u0, i0, xref_win=pr["xref_win"], rng=pr["rng"])
        ms = s.launch_timed(4, flush_l2=False)
        print(mode, "B", B, "ms", ms[1:].mean(), s.kernel_info()["ctas"], s.kernel_info()["threads_per_cta"])
        if mode == "speculative_ls" and B >= 2: break
