#!/usr/bin/env python3
"""Where the end-to-end time of one batched solve goes (host buffers): argument marshalling in Python, the C call
(pinned staging + H2D + kernel + D2H + copy-out) and the kernel itself.  python tools/e2e_breakdown.py [B] [lib]"""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sde4mbrl_px4_b200 import config, model_io, solver, synthetic  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
lib = sys.argv[2] if len(sys.argv) > 2 else None
cfgd = config.load_yaml(os.path.join(ROOT, "configs", "iris_traj.yaml"))
cfg = config.build_config(cfgd, convert_to_enu=True, rtol=0.0, atol=0.0)
s = solver.MPCSolver(cfg, model_io.synthetic_model("iris").to_blob(), lib_path=lib)
pr = synthetic.batched_problems(B, cfg.horizon, np.array(cfg.dt[: cfg.horizon]), seed=0)
u0, i0 = s.reset(B)
for _ in range(2):
    s.solve(pr["x"], u0, i0, xref_win=pr["xref_win"], rng=pr["rng"])
ta, tc, tk, tt = [], [], [], []
for _ in range(8):
    t0 = time.perf_counter()
    a, keep = s._args(pr["x"], u0, i0, None, None, pr["xref_win"], pr["rng"], None, False)
    t1 = time.perf_counter()
    s._check(s.lib.sdempc_solve_ex(s._h, C.byref(a)))
    t2 = time.perf_counter()
    ta.append((t1 - t0) * 1e3); tc.append((t2 - t1) * 1e3); tk.append(s.last_launch_ms()); tt.append((t2 - t0) * 1e3)
m = lambda v: float(np.median(v))
print(f"B={B}: python marshalling {m(ta):.3f} ms | C call {m(tc):.3f} ms of which kernel {m(tk):.3f} ms -> staging + copies "
      f"{m(tc) - m(tk):.3f} ms | total {m(tt):.3f} ms")
