#!/usr/bin/env python3
"""BASELINE config 5: Monte-Carlo closed-loop evaluation, R rollouts x T control ticks of the learned-SDE plant under
the iris trajectory MPC, sharded over the GPUs of one box (one process per GPU, no collective inside the loop, one
statistics gather at the end: sharding.closed_loop_sharded).  Prints one JSON line on rank 0.

  python tools/closed_loop_mc.py --rollouts 128 --ticks 500                       # one GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29540 \
      tools/closed_loop_mc.py --rollouts 1024 --ticks 500                         # the BASELINE shape on 8 GPUs
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sde4mbrl_px4_b200 import config, model_io, sharding, solver, synthetic, trajectory  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rollouts", type=int, default=1024)
ap.add_argument("--ticks", type=int, default=500)
ap.add_argument("--iters", type=int, default=200)
a = ap.parse_args()

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
else:                                   # same code path on one GPU: a single-rank gloo group
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29541")
    dist.init_process_group("gloo", rank=0, world_size=1)
dev = torch.device("cuda", local) if world > 1 else None

cfgd = config.load_yaml(os.path.join(ROOT, "configs", "iris_traj.yaml"))
cfg = config.build_config(cfgd, max_iter=a.iters, rtol=0.0, atol=0.0)
s = solver.MPCSolver(cfg, model_io.synthetic_model("iris").to_blob(), device=local)
tab = trajectory.csv_rows_to_table(trajectory.lemniscate(2.0, 8.0, 0.0, duration=60.0))
s.set_trajectory(tab)
R = a.rollouts
x0 = synthetic.initial_states(tab[0, 1:4], R, seed=7)
t0 = np.random.default_rng(3).uniform(0, 8, R).astype(np.float32)
x0[:, 0:3] += trajectory.interp_table(tab, t0)[:, 0:3] - tab[0, 1:4]
rng = np.array([[9000 + r, 0] for r in range(R)], np.uint64)
s.closed_loop(x0[:2], t0[:2], rng[:2], 2, want_hist=False)        # context / module warm-up
dist.barrier()
torch.cuda.synchronize()
t = time.perf_counter()
out = sharding.closed_loop_sharded(s, x0, t0, rng, a.ticks, device=dev)
torch.cuda.synchronize()
wall = time.perf_counter() - t
if rank == 0:
    st = out["stats"]
    print(json.dumps({
        "bench": "closed_loop_monte_carlo", "n_gpus": world, "rollouts": R, "ticks": a.ticks, "iterations_per_tick": a.iters,
        "rollouts_per_gpu": out["rollouts_per_rank"], "device_s_max_over_ranks": out["device_s"], "wall_s": wall,
        "ticks_per_s": out["ticks_per_s"], "rollouts_per_s": out["rollouts_per_s"],
        "rms_tracking_error_m": {"median": float(np.median(st[:, 0])), "max": float(st[:, 0].max())},
        "mean_opt_cost": float(st[:, 2].mean()), "mean_iterations": float(st[:, 3].mean()), "kernel": s.kernel_info(),
        "note": "plant = the same learned SDE with an independent Philox stream; one launch per GPU, no host sync "
                "inside the tick loop; the only exchange is the final statistics gather"}, default=float), flush=True)
dist.barrier()
dist.destroy_process_group()
