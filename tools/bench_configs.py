#!/usr/bin/env python3
"""Secondary BASELINE configs (bench.py carries the headline line):
  config 3: hexacopter, 8 particles (W = 64 synthetic model): single-tick latency and batched throughput
  config 5: Monte-Carlo closed loop, R rollouts x T ticks on device (per GPU share of 1024 x 500)
  plus iris P = 8 and the early-stopping (YAML tolerances) variant of the batched solve.
Prints one JSON object per line.  Usage: python tools/bench_configs.py [--quick]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sde4mbrl_px4_b200 import config, model_io, solver, synthetic, trajectory  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--quick", action="store_true")
ap.add_argument("--cpu", action="store_true", help="also time the CPU oracle on bounded samples")
a = ap.parse_args()


def setup(vehicle, **ov):
    cfgd = config.load_yaml(os.path.join(ROOT, "configs", f"{vehicle}_traj.yaml"))
    cfg = config.build_config(cfgd, **ov)
    return cfg, model_io.synthetic_model(vehicle).to_blob()


def pct(v, q):
    return float(np.percentile(v, q))


def tick_latency(vehicle, P, ticks, **ov):
    cfg, blob = setup(vehicle, num_particles=P, rtol=0.0, atol=0.0, **ov)
    s = solver.MPCSolver(cfg, blob)
    tab = trajectory.csv_rows_to_table(trajectory.lemniscate(2.0, 8.0, 0.0, duration=60.0))
    s.set_trajectory(tab)
    x = tab[0:1, 1:].copy()
    x[0, 0:3] += [0.3, -0.2, 0.1]
    u, i = s.reset(1)
    rng = np.array([[10, 0]], np.uint64)
    e2e, dev = [], []
    for k in range(10 + ticks):
        t = time.perf_counter()
        u, xe, i, _ = s.solve(x, u, i, curr_t=np.array([0.05 * k], np.float32), rng=rng)
        d = (time.perf_counter() - t) * 1e3
        if k >= 10:
            e2e.append(d); dev.append(i[0, 7] * 1e-3)
        x = xe[:, 1].copy(); rng[0, 1] += 1
    return {"bench": "single_tick_latency_ms", "vehicle": vehicle, "particles": P, "iterations": cfg.max_iter, "ticks": ticks,
            "e2e": {"p50": pct(e2e, 50), "p99": pct(e2e, 99), "max": max(e2e)}, "device": {"p50": pct(dev, 50), "p99": pct(dev, 99)},
            "mean_n_ls": float(i[0, 0]), "kernel": s.kernel_info()}


def batched(vehicle, P, B, launches, **ov):
    cfg, blob = setup(vehicle, num_particles=P, **ov)
    s = solver.MPCSolver(cfg, blob)
    pr = synthetic.batched_problems(B, cfg.horizon, np.array(cfg.dt[: cfg.horizon]), seed=0)
    u0, i0 = s.reset(B)
    s.stage(pr["x"], u0, i0, xref_win=pr["xref_win"], rng=pr["rng"])
    s.launch_timed(1, flush_l2=True)
    ms = s.launch_timed(launches, flush_l2=True)
    u, xe, info = s.fetch()
    return {"bench": "batched_solves_per_sec", "vehicle": vehicle, "particles": P, "B": B, "value": float(B / (ms.mean() * 1e-3)),
            "ms_per_launch": float(ms.mean()), "mean_iterations": float(info[:, 2].mean()), "mean_n_ls": float(info[:, 0].mean()),
            "early_stop": not (ov.get("rtol", 1) == 0.0), "kernel": s.kernel_info()}


def closed_loop(R, ticks, iters):
    cfg, blob = setup("iris", max_iter=iters, rtol=0.0, atol=0.0)
    s = solver.MPCSolver(cfg, blob)
    tab = trajectory.csv_rows_to_table(trajectory.lemniscate(2.0, 8.0, 0.0, duration=60.0))
    s.set_trajectory(tab)
    x0 = synthetic.initial_states(tab[0, 1:4], R, seed=7)
    t0 = np.random.default_rng(3).uniform(0, 8, R).astype(np.float32)
    x0[:, 0:3] += trajectory.interp_table(tab, t0)[:, 0:3] - tab[0, 1:4]
    rng = np.array([[9000 + r, 0] for r in range(R)], np.uint64)
    s.closed_loop(x0[:2], t0[:2], rng[:2], 2, want_hist=False)   # context / module warm-up
    t = time.perf_counter()
    _, _, st = s.closed_loop(x0, t0, rng, ticks, want_hist=False)
    wall = time.perf_counter() - t
    dev = s.last_launch_ms() * 1e-3
    return {"bench": "closed_loop_monte_carlo", "rollouts": R, "ticks": ticks, "iterations_per_tick": iters,
            "device_s": dev, "wall_s": wall, "ticks_per_s": R * ticks / dev, "rollouts_per_s": R / dev,
            "rms_tracking_error_m": {"median": float(np.median(st[:, 0])), "max": float(st[:, 0].max())},
            "mean_opt_cost": float(st[:, 2].mean()), "kernel": s.kernel_info()}


q = a.quick
jobs = [
    lambda: tick_latency("iris", 1, 50 if q else 200),
    lambda: tick_latency("hexa", 1, 20 if q else 100),
    lambda: tick_latency("hexa", 8, 20 if q else 100),
    lambda: tick_latency("iris", 8, 20 if q else 100),
    lambda: batched("hexa", 8, 148 if q else 592, 2, rtol=0.0, atol=0.0),
    lambda: batched("hexa", 1, 1184 if q else 4096, 2, rtol=0.0, atol=0.0),
    lambda: batched("iris", 8, 148 if q else 592, 2, rtol=0.0, atol=0.0),
    lambda: batched("iris", 1, 4096, 2),                       # YAML tolerances: early stopping allowed
    lambda: closed_loop(128, 20 if q else 500, 200),
    lambda: closed_loop(1024, 10 if q else 100, 200),
]
for job in jobs:
    print(json.dumps(job(), default=float), flush=True)
