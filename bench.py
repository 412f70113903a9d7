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
  e2e    : the same metric through the public API (`MPCSolver.solve_pinned` == the C ABI `sdempc_solve_ex` on the
           caller's page-locked HOST buffers: the library copies straight between them and the device): H2D of the
           step's inputs, the launch, D2H of plan + trajectory + telemetry, and for N > 1 the result gather to rank 0
           (`sharding.solve_sharded`).  `e2e.pageable_arrays_value`: the same through `MPCSolver.solve` with plain numpy
           arrays (one more host copy each way through the library's pinned staging blocks).
  latency_ms : BASELINE config 2, single iris tick (B=1) warm-started along a trajectory, p50/p99/max, both
           end to end (host call -> result in host memory, the reference's own definition of solve time,
           sde_control.py:386-425) and device only; `hexa_p8` = BASELINE config 3 as a single tick.
  strong_scaling : BASELINE config 4 as worded (`--batch` problems in TOTAL, split over the N GPUs), next to the weak
           figure (`--batch` per GPU); `--scaling strong` makes it the headline `value` instead.
  tensor_solve : the same solve on the tcgen05 mapping (SDEMPC_F_TENSOR), batched, against the FP32 kernels.
  closed_loop : BASELINE config 5 share of this job: 128 rollouts per GPU x 500 ticks of plant + MPC, one launch per GPU.
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
    # MEASURED_PEAKS.json carries no FP32-pipe figure.  Nominal: 148 SM x 128 lanes x 2 flop x max SM clock; the roofline
    # denominator is the figure the library's FFMA probe measures on this device in this run (sdempc_probe_fp32).
    out["fp32_tflops_nominal"] = 148 * 128 * 2 * out["sm_max_mhz"] * 1e6 / 1e12
    out["fp32_tflops"] = out["fp32_tflops_nominal"]
    out["fp32_source"] = f"nominal 148 SM x 128 lanes x 2 x {out['sm_max_mhz']:.0f} MHz"
    return out


def measured_peaks(device: int) -> dict:
    out = peaks()
    try:
        from sde4mbrl_px4_b200 import solver

        out["fp32_tflops"] = solver.probe_fp32_tflops(device)
        out["fp32_source"] = ("measured in this run by sdempc_probe_fp32 (FFMA-bound kernel, 16 warps/SM, CUDA events); "
                              f"nominal {out['fp32_tflops_nominal']:.1f}")
    except Exception as e:   # keep the nominal figure, say so
        out["fp32_source"] += f" (probe failed: {e})"
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
    so the CPU arm is this repo's restatement of the same solve (oracle, kind 'port') on all host threads, on the SAME
    workload as the GPU arm: every step solves the full `--batch` (per GPU for weak scaling, in total for strong)."""
    if rank != 0:
        return
    cfg, blob, pr = build_workload(args.batch, 0, args.max_iter)
    from oracle import oracle as O

    o = O.Oracle(cfg, blob, "f32")
    cores = O.set_threads(len(os.sched_getaffinity(0)))
    n = pr["x"].shape[0]
    u0, i0 = o.reset(n)
    kw = dict(xref_win=pr["xref_win"], rng=pr["rng"])
    for _ in range(max(min(args.warmup, 1), 1)):
        o.solve(pr["x"][:4 * cores], u0[:4 * cores], i0[:4 * cores], xref_win=pr["xref_win"][:4 * cores], rng=pr["rng"][:4 * cores])
    t0 = time.perf_counter()
    for _ in range(args.steps):
        o.solve(pr["x"], u0, i0, **kw)
    el = time.perf_counter() - t0
    v = n * args.steps / el
    lat = []
    for b in range(min(30, n)):   # the single solve (config 1 / 2) on one host core
        t = time.perf_counter()
        o.solve(pr["x"][b:b + 1], u0[:1], i0[:1], xref_win=pr["xref_win"][b:b + 1], rng=pr["rng"][b:b + 1])
        lat.append((time.perf_counter() - t) * 1e3)
    line = {
        "impl": "reference", "metric": "batched_mpc_solves_per_sec", "value": v, "unit": "solves/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": el / args.steps * 1e3, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(cfg, args),
        "cpu_baseline": {"value": v, "unit": "solves/s", "cores": cores, "kind": "port",
                         "sample": f"each step = all {n} iris problems of the GPU arm's step, {cfg.max_iter} iterations, oracle float32 C (-O3), "
                                   f"OpenMP over problems on {cores} threads (os.sched_getaffinity)",
                         "single_solve_ms_p50": float(np.median(lat)), "single_solve_ms_p99": float(np.percentile(lat, 99)),
                         "single_solve_threads": 1},
        "e2e": {"value": v, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "latency_ms": {"config": "iris single solve, one host thread", "e2e": {"p50": float(np.median(lat)), "p99": float(np.percentile(lat, 99))}},
        "gpu_launches": 0, "host_threads": cores,
    }
    print(json.dumps(line))


def workload_config(cfg, args, per_gpu=None) -> dict:
    return {"workload": "iris batched MPC solve (BASELINE config 4 shape): independent problems with random initial states and "
                        "lemniscate reference windows, iris_traj.yaml cost/line-search constants, synthetic learned-SDE model (W=32, L=2)",
            "problems_per_gpu_per_step": per_gpu or args.batch, "horizon": cfg.horizon, "particles": cfg.num_particles,
            "iterations": cfg.max_iter, "early_stop": False, "parallelism": f"independent problems, {args.gpus} x static shard",
            "l2": "flushed (256 MiB memset) between timed launches"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--batch", type=int, default=4096, help="problems per GPU per step (weak) / in total (strong)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="which figure is the headline `value`; the other one is reported under its own key")
    ap.add_argument("--max-iter", type=int, default=200)
    ap.add_argument("--latency-ticks", type=int, default=1000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-latency", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip tensor_solve / tensor_path / closed_loop / strong_scaling")
    ap.add_argument("--closed-loop-ticks", type=int, default=500)
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

    from sde4mbrl_px4_b200 import config, model_io, sharding, solver, synthetic, trajectory

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

    def throughput(B: int, timed_steps: int, warm: int):
        """Device-timed and end-to-end solves/s of one batched solve of B problems per GPU (all ranks)."""
        cfg, blob, pr = build_workload(B, rank, args.max_iter)
        s = solver.MPCSolver(cfg, blob, device=local)
        u0, i0 = s.reset(B)
        kw = dict(xref_win=pr["xref_win"], rng=pr["rng"])
        s.stage(pr["x"], u0, i0, **kw)
        s.launch_timed(warm, flush_l2=True)
        barrier()
        l0 = s.launch_count()
        sampler.start()
        ms = s.launch_timed(timed_steps, flush_l2=True)
        sampler.stop()
        barrier()
        launches = s.launch_count() - l0
        u, xe, info = s.fetch()
        dev_ms = max_over_ranks(float(ms.sum()))
        # end to end through the public API, host buffers
        h2d = pr["x"].nbytes + u0.nbytes + i0.nbytes + pr["xref_win"].nbytes + pr["rng"].nbytes
        d2h = u.nbytes + xe.nbytes + info.nbytes
        ue = None
        pb = None
        if dist is not None:   # N > 1: this rank's (constant) inputs page-locked in place: solve_sharded sends them by DMA
            s.pin(pr["x"], pr["xref_win"], pr["rng"])
        if dist is None:   # one GPU: the caller's buffers are page-locked (MPCSolver.pinned_batch), the library copies straight
            pb = s.pinned_batch(B, window=True)     # between them and the device: same C call (sdempc_solve_ex), no staging copy
            pb.x[:] = pr["x"]; pb.xref[:] = pr["xref_win"]; pb.rng[:] = pr["rng"]

        def one_step():
            if dist is None:
                pb.u[:] = u0; pb.info[:] = i0       # the solve works in place: this step's plan and telemetry inputs
                return s.solve_pinned(pb)[0]
            # H2D + solve + result gather to rank 0 through shared memory (no collective on the data path) + D2H
            g = sharding.solve_sharded(s, pr, u0, i0, world * B)
            return g["u"][:B] if rank == 0 else None

        for _ in range(2):
            one_step()
        barrier()
        sampler.start()
        t0 = time.perf_counter()
        for _ in range(timed_steps):
            ue = one_step()
        barrier()
        e2e_s = max_over_ranks(time.perf_counter() - t0)
        sampler.stop()
        if rank == 0:
            assert np.array_equal(ue, u), "e2e solve and staged solve disagree"
        e2e_pageable = None
        if dist is None:   # for comparison: the same through MPCSolver.solve with pageable numpy arrays (staging copies on the host)
            t0 = time.perf_counter()
            for _ in range(timed_steps):
                s.solve(pr["x"], u0, i0, **kw)
            e2e_pageable = B * timed_steps / (time.perf_counter() - t0)
            pb.close()
        return dict(cfg=cfg, blob=blob, pr=pr, s=s, info=info, B=B, value=world * B * timed_steps / (dev_ms * 1e-3),
                    ms_per_step=dev_ms / timed_steps, kern_ms=float(ms.mean()), launches=int(launches),
                    e2e_value=world * B * timed_steps / e2e_s, e2e_ms=e2e_s / timed_steps * 1e3, h2d=int(h2d), d2h=int(d2h),
                    e2e_pageable=e2e_pageable)

    B_weak, B_strong = args.batch, max(1, args.batch // world)
    main_B = B_weak if args.scaling == "weak" else B_strong
    r = throughput(main_B, args.steps, args.warmup)
    cfg, blob, pr, s, info, B = r["cfg"], r["blob"], r["pr"], r["s"], r["info"], r["B"]
    H, nu, P = cfg.horizon, cfg.nu, cfg.num_particles
    other = None
    if world > 1 and not args.no_extras:   # the other reading of config 4, fewer steps
        ro = throughput(B_strong if args.scaling == "weak" else B_weak, max(2, args.steps // 2), max(3, args.warmup))
        other = {"scaling": "strong" if args.scaling == "weak" else "weak", "value": ro["value"], "unit": "solves/s",
                 "problems_per_gpu_per_step": ro["B"], "problems_total_per_step": ro["B"] * world, "ms_per_step": ro["ms_per_step"],
                 "e2e": {"value": ro["e2e_value"], "ms_per_step": ro["e2e_ms"]}, "kernel_info": ro["s"].kernel_info(),
                 "note": "BASELINE config 4 as worded: the batch is split over the GPUs, so per-GPU work shrinks with N" if args.scaling == "weak" else
                         "per-GPU work fixed as N grows"}
        ro["s"].close()

    pk = measured_peaks(local)

    # ---------------- single-tick latency (BASELINE configs 2 and 3), rank 0 only ----------------
    def tick_latency(vehicle: str, particles: int, ticks: int, warm: int) -> dict:
        cfgd = config.load_yaml(os.path.join(ROOT, "configs", f"{vehicle}_traj.yaml"))
        c1 = config.build_config(cfgd, convert_to_enu=True, max_iter=args.max_iter, rtol=0.0, atol=0.0, num_particles=particles)
        blob1 = model_io.synthetic_model(vehicle).to_blob()
        s1 = solver.MPCSolver(c1, blob1, device=local)
        tab = trajectory.csv_rows_to_table(trajectory.lemniscate(2.0, 8.0, 0.0, duration=60.0))
        s1.set_trajectory(tab)
        x = tab[0:1, 1:].copy()
        x[0, 0:3] += [0.3, -0.2, 0.1]
        up, ip = s1.reset(1)
        rng1 = np.array([[10, 0]], np.uint64)
        e2e_l, dev_l, flop_l, it_l = [], [], [], []
        for k in range(warm + ticks):
            ct = np.array([0.05 * k], np.float32)
            t = time.perf_counter()
            up, xe1, ip, _ = s1.solve(x, up, ip, curr_t=ct, rng=rng1)
            dtm = (time.perf_counter() - t) * 1e3
            if k >= warm:
                e2e_l.append(dtm)
                dev_l.append(float(ip[0, 7]) * 1e-3)
                flop_l.append(algorithmic_flops(ip, c1.horizon, particles, F_STEP[vehicle]))
                it_l.append(float(ip[0, 2]))
            x = xe1[:, 1].copy()
            rng1[0, 1] += 1
        pc = lambda a, q: float(np.percentile(a, q))
        out = {"config": f"{vehicle} single tick, B=1, P={particles}, H={c1.horizon}, {args.max_iter} iterations, warm-started along a lemniscate",
               "ticks": ticks, "e2e": {"p50": pc(e2e_l, 50), "p99": pc(e2e_l, 99), "max": float(np.max(e2e_l))},
               "device": {"p50": pc(dev_l, 50), "p99": pc(dev_l, 99), "max": float(np.max(dev_l))}}
        # the tick as a fraction of the FP32 roofline (north_star) and its dependent chain: a tick is one
        # problem on the SMs of one cluster, so the whole-GPU fraction is tiny by construction; the chain length is
        # the number that matters: every iteration is H forward + H adjoint step evaluations in sequence
        ki1 = s1.kernel_info()
        ach = float(np.mean(flop_l)) / (pc(dev_l, 50) * 1e-3) / 1e12
        sms = max(1, min(int(ki1["ctas"]), int(ki1["sm_count"])))
        out["roofline"] = {
            "bound": "latency (dependent chain)", "algorithmic_flop_per_tick": float(np.mean(flop_l)),
            "achieved_tflops": ach, "frac_of_gpu_fp32_peak": ach / pk["fp32_tflops"], "sms_occupied": sms,
            "frac_of_occupied_sms_fp32_peak": ach / (pk["fp32_tflops"] * sms / int(ki1["sm_count"])),
            "cycles_per_step_evaluation_on_the_critical_path":
                pc(dev_l, 50) * 1e-3 * pk["sm_max_mhz"] * 1e6 / (float(np.mean(it_l)) * 2 * c1.horizon),
            "chain_note": "tick cycles / (iterations x 2H): an upper bound — an iteration whose outcome was not speculated "
                          "(5-12 % of them) adds a line-search pass and a gradient pass to the chain",
            "kernel": ki1}
        if not args.no_cpu_baseline:   # the CPU statement of the same tick on one host core
            from oracle import oracle as O

            o1 = O.Oracle(c1, blob1, "f32")
            o1.set_trajectory(tab)
            O.set_threads(1)
            upo, ipo = o1.reset(1)
            xo, cl = tab[0:1, 1:].copy(), []
            for k in range(12):
                t = time.perf_counter()
                upo, xeo, ipo, _ = o1.solve(xo, upo, ipo, curr_t=np.array([0.05 * k], np.float32), rng=np.array([[10, k]], np.uint64))
                cl.append((time.perf_counter() - t) * 1e3)
                xo = xeo[:, 1].copy()
            out["cpu_port_single_thread_ms_p50"] = float(np.median(cl[2:]))
        s1.close()
        return out

    latency = None
    if rank == 0 and not args.no_latency:
        latency = tick_latency("iris", 1, args.latency_ticks, 100)
        latency["hexa_p8"] = tick_latency("hexa", 8, max(50, args.latency_ticks // 4), 20)

    # ---------------- the solve on the tcgen05 mapping (SDEMPC_F_TENSOR), rank 0 of a single-GPU run ----------------
    tensor_solve = tensor_path = None
    if rank == 0 and world == 1 and not args.no_extras:
        def solve_pair(vehicle, particles, Bt, fp32_too=True):
            cfgd = config.load_yaml(os.path.join(ROOT, "configs", f"{vehicle}_traj.yaml"))
            blob_v = model_io.synthetic_model(vehicle).to_blob()
            ov = dict(convert_to_enu=True, num_particles=particles, max_iter=args.max_iter, rtol=0.0, atol=0.0)
            Hv = int(cfgd["horizon"])
            out = {"workload": f"{Bt} {vehicle} problems x {particles} particle(s) = {Bt * particles} rollout rows, {args.max_iter} iterations, sdempc_solve_ex"}
            res = {}
            for name, c in (("tcgen05_tf32", config.build_config(cfgd, tensor=True, **ov)),) + ((("fp32", config.build_config(cfgd, **ov)),) if fp32_too else ()):
                prt = synthetic.batched_problems(Bt, Hv, np.array(c.dt[:Hv]), seed=7)
                st = solver.MPCSolver(c, blob_v, device=local)
                u0t, i0t = st.reset(Bt)
                st.stage(prt["x"], u0t, i0t, xref_win=prt["xref_win"], rng=prt["rng"])
                ms = st.launch_timed(3, flush_l2=True)
                _, _, inf = st.fetch()
                ki = st.kernel_info()
                res[name] = inf
                msv = float(ms[1:].mean())
                flop = algorithmic_flops(inf, Hv, particles, F_STEP[vehicle])
                out[name] = {"ms": msv, "solves_per_sec": Bt / msv * 1e3, "algorithmic_tflops": flop / (msv * 1e-3) / 1e12,
                             "ctas": ki["ctas"], "problems_per_cta": ki["problems_per_cta"], "mean_linesearch_trials": float(inf[:, 0].mean())}
                st.close()
            if fp32_too:
                rc = np.abs(res["tcgen05_tf32"][:, 6] - res["fp32"][:, 6]) / np.abs(res["fp32"][:, 6])
                out["speedup_vs_fp32"] = out["fp32"]["ms"] / out["tcgen05_tf32"]["ms"]
                out["opt_cost_rel_diff_vs_fp32"] = {"median": float(np.median(rc)), "p90": float(np.quantile(rc, 0.9)), "max": float(rc.max())}
            return out

        tensor_solve = {"note": "supplementary: NOT the headline (TF32 products; compared with the oracle teacher-forced and at cost level, "
                                "tests/test_gpu_parity.py::test_tensor_core_solve_*); FP32 = the bit-exact kernels the headline uses",
                        "iris_4096": solve_pair("iris", 1, 4096), "iris_65536": solve_pair("iris", 1, 65536),
                        "hexa_p8_8192": solve_pair("hexa", 8, 8192)}

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
                for _ in range(6):
                    Jr = sr.rollout(prt["x"], ut, upt, xref_win=prt["xref_win"], rng=prt["rng"], want_grad=False)[0]
                    ms.append(sr.last_launch_ms())
                msg = []
                for _ in range(5):
                    gr = sr.rollout(prt["x"], ut, upt, xref_win=prt["xref_win"], rng=prt["rng"], want_grad=True)[1]
                    msg.append(sr.last_launch_ms())
                res[name] = (float(np.median(ms[2:])), Jr, float(np.median(msg[2:])), gr)   # warm-up launches dropped
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
        tensor_path["hexa_8_particles"] = tc_pair("hexa", 8, 8192)

    # ---------------- Monte-Carlo closed loop (BASELINE config 5): 128 rollouts per GPU x 500 ticks ----------------
    closed_loop = None
    if not args.no_extras:
        R = 128 * world
        tab = trajectory.csv_rows_to_table(trajectory.lemniscate(2.0, 8.0, 0.0, duration=60.0))
        cfgd = config.load_yaml(os.path.join(ROOT, "configs", "iris_traj.yaml"))
        ccl = config.build_config(cfgd, convert_to_enu=True, max_iter=args.max_iter, rtol=0.0, atol=0.0)
        scl = solver.MPCSolver(ccl, blob, device=local)
        scl.set_trajectory(tab)
        x0 = synthetic.initial_states(tab[0, 1:4], R, seed=7)
        t0s = np.random.default_rng(3).uniform(0, 8, R).astype(np.float32)
        x0[:, 0:3] += trajectory.interp_table(tab, t0s)[:, 0:3] - tab[0, 1:4]
        rngs = np.array([[9000 + q, 0] for q in range(R)], np.uint64)
        barrier()
        if dist is None:
            _, _, st = scl.closed_loop(x0, t0s, rngs, args.closed_loop_ticks, want_hist=False)
            dev_s = scl.last_launch_ms() * 1e-3
            res = {"stats": st, "device_s": dev_s, "rollouts_per_rank": [R]}
        else:
            res = sharding.closed_loop_sharded(scl, x0, t0s, rngs, args.closed_loop_ticks, device=torch.device("cuda", local))
        barrier()
        if rank == 0:
            st = res["stats"]
            closed_loop = {"workload": f"{R} rollouts x {args.closed_loop_ticks} control ticks of learned-SDE plant + iris trajectory MPC "
                                       f"({args.max_iter} iterations per tick), {R // world} rollouts per GPU, one launch per GPU, no host sync "
                                       "inside the tick loop (BASELINE config 5 is 1024 x 500 on 8 GPUs = this at N = 8)",
                           "device_s": res["device_s"], "ticks_per_sec": R * args.closed_loop_ticks / res["device_s"],
                           "rollouts_per_sec": R / res["device_s"], "rollouts_per_gpu": res["rollouts_per_rank"],
                           "rms_tracking_error_m": {"median": float(np.median(st[:, 0])), "max": float(st[:, 0].max())},
                           "mean_iterations_per_tick": float(st[:, 3].mean())}
        scl.close()

    # ---------------- roofline of the dominant (only) kernel ----------------
    ki = s.kernel_info()
    kernel_name = ("mpc_group_kernel<4,32,GP=4,GW=8>" if ki["problems_per_cta"] == 32 else
                   "mpc_kernel<4,32,1,G=%d,SOLVE>" % ki["problems_per_cta"])
    kern_ms = r["kern_ms"]
    flops_launch = algorithmic_flops(info, H, P, F_STEP["iris"])
    achieved_tf = flops_launch / (kern_ms * 1e-3) / 1e12
    alg_bytes = B * (13 * 4 + (H + 1) * 13 * 4 + 2 * H * nu * 4 + (H + 1) * 13 * 4 + 2 * 32 + 16) + 2 * (1536 + 70) * 4
    traffic = None   # dram__bytes_read.sum + dram__bytes_write.sum per launch, from the committed ncu --set full capture
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp) and B == 4096 and args.max_iter == 200:
        traffic = json.load(open(tp)).get("dram_bytes_per_launch")
    roofline = {"bound": "fp32", "achieved": achieved_tf, "peak": pk["fp32_tflops"], "unit": "TFLOP/s",
                "frac": achieved_tf / pk["fp32_tflops"], "traffic": traffic,
                "peak_source": pk["fp32_source"], "frac_of_nominal_peak": achieved_tf / pk["fp32_tflops_nominal"],
                "kernel": kernel_name, "avg_launch_ms": kern_ms, "algorithmic_flop_per_launch": flops_launch,
                "hbm": {"algorithmic_bytes_per_launch": alg_bytes, "achieved_gbs": alg_bytes / (kern_ms * 1e-3) / 1e9,
                        "peak_gbs": pk["hbm_gbs"], "note": "not HBM bound: ~3 KB per problem"},
                "mean_linesearch_trials": float(info[:, 0].mean())}

    cb = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cb = cpu_baseline(cfg, blob, pr)

    if rank == 0:
        line = {
            "metric": "batched_mpc_solves_per_sec", "value": r["value"], "unit": "solves/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(cfg, args, per_gpu=B),
            "e2e": {"value": r["e2e_value"], "unit": "solves/s", "h2d_bytes_per_step": r["h2d"], "d2h_bytes_per_step": r["d2h"],
                    "ms_per_step": r["e2e_ms"],
                    "api": "MPCSolver.solve_pinned: sdempc_solve_ex on page-locked caller buffers (N = 1); "
                           "sharding.solve_sharded (N > 1)",
                    "pageable_arrays_value": r["e2e_pageable"]},
            "gpu_launches": r["launches"], "gpu_launches_e2e": int(args.steps),
            "roofline": roofline, "cpu_baseline": cb, "latency_ms": latency, "clocks": sampler.summary(),
            "kernel_info": ki, ("strong_scaling" if args.scaling == "weak" else "weak_scaling"): other,
            "tensor_solve": tensor_solve, "tensor_path": tensor_path, "closed_loop": closed_loop,
        }
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
