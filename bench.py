#!/usr/bin/env python3
"""Benchmark of the MPC-solve hot path (BASELINE.json metric: single-solve p50/p99 latency and batched
solves/s at 1/2/4/8 B200, each next to the FP32 roofline and the CPU timing of the same solve).

    python bench.py --gpus N --steps K --warmup W            # this framework (one rank per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # the CPU statement of the same solve on the host cores

One "step" = one batched solve of `--batch` independent iris MPC problems per GPU (BASELINE config 4: 4096
random initial states / lemniscate reference windows, H=20, P=1, 200 APG iterations with rtol=atol=0 so that
iteration counts are identical everywhere).  Per-GPU work is fixed as N grows: weak scaling, no collective
on the data path; the only exchange is the final result gather, which is inside the e2e region.
  value  : solves/s with inputs resident in HBM, CUDA-event time of the K launches on the launching stream,
           L2 flushed between launches, max over ranks.
  e2e    : the same metric through the public API (`MPCSolver.solve` == the C ABI `sdempc_solve_ex`) with HOST
           buffers: H2D of the step's inputs from pinned staging, the launch, D2H of plan + trajectory +
           telemetry, and for N > 1 the result gather to rank 0.
  latency_ms : BASELINE config 2, single iris tick (B=1) warm-started along a trajectory, p50/p99/max, both
           end to end (host call -> result in host memory, the reference's own definition of solve time,
           sde_control.py:386-425) and device only.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

F_STEP = {"iris": 6444.0, "hexa": 21292.0}   # algorithmic flop per EM step forward


def algorithmic_flops(info: np.ndarray, H: int, P: int, f_step: float) -> float:
    """F_solve summed over problems: per iteration 1 forward + adjoint (= 2 forwards) + n_ls forward trials."""
    iters = info[:, 2].astype(np.float64)
    n_ls_total = info[:, 0].astype(np.float64) * iters
    return float(np.sum(H * P * f_step * (3.0 * iters + n_ls_total)))


def peaks() -> dict:
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    out = {"sm_max_mhz": 1965.0, "hbm_gbs": 6650.0, "source": "fallback"}
    if os.path.exists(p):
        try:
            m = json.load(open(p))
            out.update(sm_max_mhz=float(m.get("sm_max_mhz", 1965.0)), hbm_gbs=float(m.get("hbm_gbs", 6650.0)), source="measured")
        except Exception:
            pass
    # MEASURED_PEAKS.json carries no FP32-pipe figure: peak = 148 SM x 128 lanes x 2 flop x max SM clock
    out["fp32_tflops"] = 148 * 128 * 2 * out["sm_max_mhz"] * 1e6 / 1e12
    return out


class ClockSampler:
    """Samples SM clock and throttle reasons during the timed regions (NVML; nvidia-smi columns of the recipe)."""

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz, self._stop, self._t = [], set(), None, threading.Event(), None
        try:
            import pynvml as nv

            nv.nvmlInit()
            self.nv, self.h = nv, nv.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.1)

    def start(self):
        if self.nv is not None:
            self._stop.clear()
            self._t = threading.Thread(target=self._loop, daemon=True)
            self._t.start()

    def stop(self):
        if self._t is not None:
            self._stop.set()
            self._t.join()
            self._t = None

    def summary(self) -> dict:
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def build_workload(batch: int, rank: int, max_iter: int):
    from sde4mbrl_px4_b200 import config, model_io, synthetic

    cfgd = config.load_yaml(os.path.join(ROOT, "configs", "iris_traj.yaml"))
    cfg = config.build_config(cfgd, convert_to_enu=True, max_iter=max_iter, rtol=0.0, atol=0.0)
    blob = model_io.synthetic_model("iris").to_blob()
    pr = synthetic.batched_problems(batch, cfg.horizon, np.array(cfg.dt[: cfg.horizon]), seed=1000 * rank)
    pr["rng"][:, 0] += np.uint64(1_000_000 * rank)
    return cfg, blob, pr


def cpu_baseline(cfg, blob, pr, budget_s: float = 12.0, chunk: int = 256) -> dict:
    """The oracle (float32, OpenMP over problems, all host threads) on a bounded sample of the same workload."""
    from oracle import oracle as O

    o = O.Oracle(cfg, blob, "f32")
    cores = O.set_threads(len(os.sched_getaffinity(0)))
    B = pr["x"].shape[0]
    u0, i0 = o.reset(min(chunk, B))
    n = min(chunk, B)
    o.solve(pr["x"][:8], u0[:8], i0[:8], xref_win=pr["xref_win"][:8], rng=pr["rng"][:8])   # warm-up
    done, t0 = 0, time.perf_counter()
    while True:
        lo = done % max(1, B - n + 1)
        o.solve(pr["x"][lo:lo + n], u0, i0, xref_win=pr["xref_win"][lo:lo + n], rng=pr["rng"][lo:lo + n])
        done += n
        el = time.perf_counter() - t0
        if el >= budget_s or done >= 4 * B:
            break
    # single-solve latency on the same inputs (B = 1: one OpenMP iteration, i.e. one thread)
    lat = []
    for b in range(min(20, B)):
        t = time.perf_counter()
        o.solve(pr["x"][b:b + 1], u0[:1], i0[:1], xref_win=pr["xref_win"][b:b + 1], rng=pr["rng"][b:b + 1])
        lat.append((time.perf_counter() - t) * 1e3)
    return {"value": done / el, "unit": "solves/s", "cores": cores, "kind": "port",
            "sample": f"{done} iris solves x {cfg.max_iter} iterations in {el:.1f} s, oracle float32 C (-O3), OpenMP over problems on {cores} threads",
            "single_solve_ms_p50": float(np.median(lat))}


def run_reference(args, rank: int, world: int):
    """--impl reference: the reference's solver source is not in /root/reference (un-vendored JAX package),
    so the CPU arm is this repo's restatement of the same solve (oracle, kind 'port') on all host threads."""
    if rank != 0:
        return
    cfg, blob, pr = build_workload(min(args.batch, 512), 0, args.max_iter)
    from oracle import oracle as O

    o = O.Oracle(cfg, blob, "f32")
    cores = O.set_threads(len(os.sched_getaffinity(0)))
    n = pr["x"].shape[0]
    u0, i0 = o.reset(n)
    kw = dict(xref_win=pr["xref_win"], rng=pr["rng"])
    for _ in range(max(args.warmup, 1)):
        o.solve(pr["x"][:64], u0[:64], i0[:64], xref_win=pr["xref_win"][:64], rng=pr["rng"][:64])
    t0 = time.perf_counter()
    for _ in range(args.steps):
        o.solve(pr["x"], u0, i0, **kw)
    el = time.perf_counter() - t0
    v = n * args.steps / el
    line = {
        "impl": "reference", "metric": "batched_mpc_solves_per_sec", "value": v, "unit": "solves/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": el / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(cfg, args, per_step=n),
        "cpu_baseline": {"value": v, "unit": "solves/s", "cores": cores, "kind": "port",
                         "sample": f"each step = {n} of the {args.batch} iris problems, {cfg.max_iter} iterations, {cores} OpenMP threads"},
        "e2e": {"value": v, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(cfg, args, per_step=None) -> dict:
    return {"workload": "iris batched MPC solve (BASELINE config 4 shape): independent problems with random initial states and "
                        "lemniscate reference windows, iris_traj.yaml cost/line-search constants, synthetic learned-SDE model (W=32, L=2)",
            "problems_per_gpu_per_step": per_step or args.batch, "horizon": cfg.horizon, "particles": cfg.num_particles,
            "iterations": cfg.max_iter, "early_stop": False, "parallelism": f"independent problems, {args.gpus} x static shard",
            "l2": "flushed (256 MiB memset) between timed launches"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--batch", type=int, default=4096, help="problems per GPU per step")
    ap.add_argument("--max-iter", type=int, default=200)
    ap.add_argument("--latency-ticks", type=int, default=1000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-latency", action="store_true")
    args = ap.parse_args()

    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the MPC solve has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"   # keep stdout to the one JSON line
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from sde4mbrl_px4_b200 import sharding, solver

    cfg, blob, pr = build_workload(args.batch, rank, args.max_iter)
    B, H, nu, P = args.batch, cfg.horizon, cfg.nu, cfg.num_particles
    s = solver.MPCSolver(cfg, blob, device=local)
    u0, i0 = s.reset(B)
    kw = dict(xref_win=pr["xref_win"], rng=pr["rng"])

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v: float) -> float:
        if dist is None:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sampler = ClockSampler(local)
    # ---------------- device-resident throughput (value) ----------------
    s.stage(pr["x"], u0, i0, **kw)
    s.launch_timed(args.warmup, flush_l2=True)
    barrier()
    l0 = s.launch_count()
    sampler.start()
    ms = s.launch_timed(args.steps, flush_l2=True)
    sampler.stop()
    barrier()
    launches = s.launch_count() - l0
    u, xe, info = s.fetch()
    dev_ms = max_over_ranks(float(ms.sum()))
    value = world * B * args.steps / (dev_ms * 1e-3)
    kern_ms = float(ms.mean())

    # ---------------- end to end through the public API, host buffers ----------------
    h2d = pr["x"].nbytes + u0.nbytes + i0.nbytes + pr["xref_win"].nbytes + pr["rng"].nbytes
    d2h = u.nbytes + xe.nbytes + info.nbytes
    for _ in range(2):
        if dist is None:
            s.solve(pr["x"], u0, i0, **kw)
        else:
            sharding.solve_sharded(s, pr, u0, i0, world * B)
    barrier()
    sampler.start()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        if dist is None:
            ue, xee, infoe, _ = s.solve(pr["x"], u0, i0, **kw)
        else:   # H2D + solve + device-to-device result gather to rank 0 (the only collective) + one D2H there
            g = sharding.solve_sharded(s, pr, u0, i0, world * B)
            if rank == 0:
                ue = g["u"][:B]
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    sampler.stop()
    launches_e2e = args.steps
    e2e_value = world * B * args.steps / e2e_s
    if rank == 0:
        assert np.array_equal(ue, u), "e2e solve and staged solve disagree"

    # ---------------- single-tick latency (BASELINE config 2), rank 0 only ----------------
    latency = None
    if rank == 0 and not args.no_latency:
        from sde4mbrl_px4_b200 import trajectory

        s1 = solver.MPCSolver(cfg, blob, device=local)
        tab = trajectory.csv_rows_to_table(trajectory.lemniscate(2.0, 8.0, 0.0, duration=60.0))
        s1.set_trajectory(tab)
        x = tab[0:1, 1:].copy()
        x[0, 0:3] += [0.3, -0.2, 0.1]
        up, ip = s1.reset(1)
        rng1 = np.array([[10, 0]], np.uint64)
        e2e_l, dev_l, flop_l, it_l = [], [], [], []
        n_warm = 100
        for k in range(n_warm + args.latency_ticks):
            ct = np.array([0.05 * k], np.float32)
            t = time.perf_counter()
            up, xe1, ip, _ = s1.solve(x, up, ip, curr_t=ct, rng=rng1)
            dtm = (time.perf_counter() - t) * 1e3
            if k >= n_warm:
                e2e_l.append(dtm)
                dev_l.append(float(ip[0, 7]) * 1e-3)
                flop_l.append(algorithmic_flops(ip, H, P, F_STEP["iris"]))
                it_l.append(float(ip[0, 2]))
            x = xe1[:, 1].copy()
            rng1[0, 1] += 1
        pc = lambda a, q: float(np.percentile(a, q))
        latency = {"config": "iris single tick, B=1, P=1, H=20, 200 iterations, warm-started along a lemniscate",
                   "ticks": args.latency_ticks,
                   "e2e": {"p50": pc(e2e_l, 50), "p99": pc(e2e_l, 99), "max": float(np.max(e2e_l))},
                   "device": {"p50": pc(dev_l, 50), "p99": pc(dev_l, 99), "max": float(np.max(dev_l))}}
        # the tick as a fraction of the FP32 roofline (north_star) and its dependent chain: a tick is one
        # problem on the SMs of one cluster, so the whole-GPU fraction is tiny by construction; the chain length is
        # the number that matters: every iteration is H forward + H adjoint step evaluations in sequence
        pkl, ki1 = peaks(), s1.kernel_info()
        ach = float(np.mean(flop_l)) / (pc(dev_l, 50) * 1e-3) / 1e12
        sms = max(1, min(int(ki1["ctas"]), int(ki1["sm_count"])))
        latency["roofline"] = {
            "bound": "latency (dependent chain)", "algorithmic_flop_per_tick": float(np.mean(flop_l)),
            "achieved_tflops": ach, "frac_of_gpu_fp32_peak": ach / pkl["fp32_tflops"], "sms_occupied": sms,
            "frac_of_occupied_sms_fp32_peak": ach / (pkl["fp32_tflops"] * sms / int(ki1["sm_count"])),
            "cycles_per_step_evaluation_on_the_critical_path":
                pc(dev_l, 50) * 1e-3 * pkl["sm_max_mhz"] * 1e6 / (float(np.mean(it_l)) * 2 * H),
            "kernel": ki1}
        s1.close()

    # ---------------- tensor-core cost evaluation (supplementary: not the headline metric), rank 0 only ----------------
    tensor_path = None
    if rank == 0 and world == 1 and not args.no_latency:
        from sde4mbrl_px4_b200 import config, model_io, synthetic

        def tc_pair(vehicle, particles, Bt):
            """sdempc_rollout on the FP32 path and on the tcgen05 path, same inputs: launch times and differences."""
            cfgd = config.load_yaml(os.path.join(ROOT, "configs", f"{vehicle}_traj.yaml"))
            blob_v = model_io.synthetic_model(vehicle).to_blob()
            cfgs = {"fp32": config.build_config(cfgd, convert_to_enu=True, num_particles=particles),
                    "tcgen05_tf32": config.build_config(cfgd, convert_to_enu=True, num_particles=particles, tensor=True)}
            Hv, nuv = cfgs["fp32"].horizon, cfgs["fp32"].nu
            prt = synthetic.batched_problems(Bt, Hv, np.array(cfgs["fp32"].dt[:Hv]), seed=7)
            ut = np.full((Bt, Hv, nuv), float(cfgs["fp32"].uref[0]), np.float32)
            upt = ut[:, 0].copy()
            res = {}
            for name, c in cfgs.items():
                sr = solver.MPCSolver(c, blob_v, device=local)
                ms = []
                for _ in range(8):
                    Jr = sr.rollout(prt["x"], ut, upt, xref_win=prt["xref_win"], rng=prt["rng"], want_grad=False)[0]
                    ms.append(sr.last_launch_ms())
                msg = []
                for _ in range(6):
                    gr = sr.rollout(prt["x"], ut, upt, xref_win=prt["xref_win"], rng=prt["rng"], want_grad=True)[1]
                    msg.append(sr.last_launch_ms())
                res[name] = (float(np.median(ms[3:])), Jr, float(np.median(msg[2:])), gr)   # warm-up launches dropped
                sr.close()
            f, t = res["fp32"], res["tcgen05_tf32"]
            rows = Bt * particles
            return {"workload": f"{Bt} problems x {particles} particle(s) x {Hv} steps (sdempc_rollout: cost evaluation, and "
                                f"value_and_grad; {vehicle})",
                    "fp32_ms": f[0], "tcgen05_tf32_ms": t[0], "tcgen05_rollouts_per_sec": rows / t[0] * 1e3,
                    "fp32_rollouts_per_sec": rows / f[0] * 1e3,
                    "max_rel_cost_error": float(np.max(np.abs(t[1] - f[1]) / np.abs(f[1]))),
                    "value_and_grad": {"fp32_ms": f[2], "tcgen05_tf32_ms": t[2],
                                       "max_grad_error_over_max_grad": float(np.max(
                                           np.abs(t[3] - f[3]).reshape(Bt, -1).max(axis=1) / np.abs(f[3]).reshape(Bt, -1).max(axis=1)))}}

        tensor_path = tc_pair("iris", 1, 65536)
        # BASELINE config 3 shape (hexacopter, 8 particles) as a batch: the same 65 536 rows
        tensor_path["hexa_8_particles"] = tc_pair("hexa", 8, 8192)

    # ---------------- roofline of the dominant (only) kernel ----------------
    ki = s.kernel_info()
    kernel_name = ("mpc_group_kernel<4,32,GP=4,GW=8>" if ki["problems_per_cta"] == 32 else
                   "mpc_kernel<4,32,1,G=%d,SOLVE>" % ki["problems_per_cta"])
    pk = peaks()
    flops_launch = algorithmic_flops(info, H, P, F_STEP["iris"])
    achieved_tf = flops_launch / (kern_ms * 1e-3) / 1e12
    alg_bytes = B * (13 * 4 + (H + 1) * 13 * 4 + 2 * H * nu * 4 + (H + 1) * 13 * 4 + 2 * 32 + 16) + 2 * (1536 + 70) * 4
    traffic = None   # dram__bytes_read.sum + dram__bytes_write.sum per launch, from the committed ncu --set full capture
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp) and B == 4096 and args.max_iter == 200:
        traffic = json.load(open(tp)).get("dram_bytes_per_launch")
    roofline = {"bound": "fp32", "achieved": achieved_tf, "peak": pk["fp32_tflops"], "unit": "TFLOP/s",
                "frac": achieved_tf / pk["fp32_tflops"], "traffic": traffic,
                "peak_source": f"148 SM x 128 lanes x 2 x {pk['sm_max_mhz']:.0f} MHz ({pk['source']} clock; MEASURED_PEAKS.json has no FP32-pipe figure)",
                "kernel": kernel_name, "avg_launch_ms": kern_ms, "algorithmic_flop_per_launch": flops_launch,
                "hbm": {"algorithmic_bytes_per_launch": alg_bytes, "achieved_gbs": alg_bytes / (kern_ms * 1e-3) / 1e9,
                        "peak_gbs": pk["hbm_gbs"], "note": "not HBM bound: ~3 KB per problem"},
                "mean_linesearch_trials": float(info[:, 0].mean())}

    cb = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cb = cpu_baseline(cfg, blob, pr)

    if rank == 0:
        line = {
            "metric": "batched_mpc_solves_per_sec", "value": value, "unit": "solves/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(cfg, args),
            "e2e": {"value": e2e_value, "unit": "solves/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": e2e_s / args.steps * 1e3},
            "gpu_launches": int(launches), "gpu_launches_e2e": int(launches_e2e),
            "roofline": roofline, "cpu_baseline": cb, "latency_ms": latency, "clocks": sampler.summary(),
            "kernel_info": s.kernel_info(), "tensor_path": tensor_path,
        }
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
