#!/usr/bin/env python3
"""Tiny solves through every kernel variant, for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sde4mbrl_px4_b200 import config, model_io, solver, synthetic, trajectory
which = sys.argv[1:] or ["cluster", "spec8", "warp", "group", "pcluster", "pcluster8", "team", "closed", "rollout", "hexa", "tcs", "tcs_spec", "tcs_p8"]
def run(name, vehicle="iris", B=3, P=1, iters=2, **flags):
    cfgd = config.load_yaml(os.path.join(ROOT, "configs", f"{vehicle}_traj.yaml"))
    cfg = config.build_config(cfgd, max_iter=iters, num_particles=P, **flags)
    s = solver.MPCSolver(cfg, model_io.synthetic_model(vehicle).to_blob())
    pr = synthetic.batched_problems(B, cfg.horizon, np.array(cfg.dt[: cfg.horizon]), seed=0)
    u0, i0 = s.reset(B)
    if name == "closed":
        tab = trajectory.csv_rows_to_table(trajectory.lemniscate(duration=5.0))
        s.set_trajectory(tab)
        s.closed_loop(pr["x"], np.zeros(B, np.float32), pr["rng"], 2)
    elif name == "rollout":
        s.rollout(pr["x"], u0, u0[:, 0], xref_win=pr["xref_win"], rng=pr["rng"])
    else:
        s.solve(pr["x"], u0, i0, xref_win=pr["xref_win"], rng=pr["rng"])
    print(name, "ok", s.kernel_info(), flush=True)
if "cluster" in which: run("cluster")
if "spec8" in which: run("spec8", no_cluster=True)
if "warp" in which: run("warp", sequential_ls=True, B=9)
if "group" in which: run("group", group=True, B=7)
if "pcluster" in which: run("pcluster", P=2, B=2)
if "pcluster8" in which: run("pcluster8", vehicle="hexa", P=8, B=1, iters=4)      # 16-CTA cluster of two warps per CTA
if "tcs" in which: run("tcs", B=148 * 20, iters=3, tensor=True)                     # tensor-core solve, plain build (20 problems per CTA)
if "tcs_spec" in which: run("tcs_spec", B=40, iters=4, tensor=True)                 # speculative build, one problem per CTA
if "tcs_p8" in which: run("tcs_p8", vehicle="hexa", P=8, B=700, iters=2, tensor=True)   # 5 problems x 8 particles per CTA: plain build
if "team" in which: run("team", P=2, B=3, sequential_ls=True)
if "closed" in which: run("closed", B=2)
if "rollout" in which: run("rollout", P=2, B=3)
if "hexa" in which: run("hexa", vehicle="hexa", B=2)
