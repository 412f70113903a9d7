"""Static sharding of independent MPC problems across ranks.

One process per GPU; problem index range [0, B) is split into contiguous blocks, each
rank solves its block with no collective on the data path, and the only exchange is the
final result gather (plan + predicted trajectory + telemetry, about 1.5 KB per problem).
``torch.distributed`` is plumbing only (NCCL on GPU boxes, gloo in CPU tests).
"""
from __future__ import annotations

import numpy as np


def shard_range(B: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block of rank ``rank``: sizes differ by at most one, earlier ranks get the extra."""
    base, rem = divmod(B, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_problem(problem: dict, rank: int, world: int) -> dict:
    B = next(iter(problem.values())).shape[0]
    lo, hi = shard_range(B, rank, world)
    return {k: v[lo:hi] for k, v in problem.items()}


def gather_results(local: dict, B: int, device=None) -> dict | None:
    """All ranks call this; rank 0 receives the concatenated [B, ...] float32 arrays, the others None.
    Uses one all_gather of a padded flat float32 buffer (NCCL has no gatherv)."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(), dist.get_rank()
    keys = sorted(local.keys())
    flat = np.concatenate([np.ascontiguousarray(local[k], np.float32).reshape(local[k].shape[0], -1) for k in keys], axis=1)
    widths = [int(np.prod(local[k].shape[1:])) for k in keys]
    shapes = [local[k].shape[1:] for k in keys]
    max_rows = max(shard_range(B, r, world)[1] - shard_range(B, r, world)[0] for r in range(world))
    pad = np.zeros((max_rows, flat.shape[1]), np.float32)
    pad[: flat.shape[0]] = flat
    t = torch.from_numpy(pad)
    if device is not None:
        t = t.to(device)
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    if rank != 0:
        return None
    rows = []
    for r in range(world):
        lo, hi = shard_range(B, r, world)
        rows.append(out[r][: hi - lo].cpu().numpy())
    allf = np.concatenate(rows, axis=0)
    res, off = {}, 0
    for k, w, s in zip(keys, widths, shapes):
        res[k] = allf[:, off: off + w].reshape((B,) + tuple(s))
        off += w
    return res


class _DevBuf:
    """Zero-copy view of a raw device allocation for torch (``__cuda_array_interface__``)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes // 4,), "typestr": "<f4", "data": (ptr, False), "version": 3,
                                         "strides": None}


_gather_cache: dict = {}


def gather_to_rank0(parts: dict, bufs: dict | None = None) -> dict | None:
    """True gather (rank 0 is the only receiver) of equally shaped per-rank tensors: ``parts[name]`` is this rank's
    flat float32 tensor (CPU for gloo, CUDA for NCCL).  Returns ``{name: [world, n] tensor}`` on rank 0 (in pinned
    host memory when the inputs are CUDA tensors), None elsewhere.  For CUDA inputs the device-to-host copy of
    array k runs on a side stream while array k + 1 is still being gathered.  ``bufs`` caches the receive / pinned
    buffers between calls."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(), dist.get_rank()
    bufs = {} if bufs is None else bufs
    names = sorted(parts.keys())
    cuda = parts[names[0]].is_cuda
    if cuda and "_copy_stream" not in bufs:
        bufs["_copy_stream"] = torch.cuda.Stream()
    out = {}
    events = []
    for k in names:
        t = parts[k]
        n = t.numel()
        if rank == 0:
            if k not in bufs or bufs[k][0].shape != (world, n):
                recv = torch.empty((world, n), dtype=torch.float32, device=t.device)
                host = torch.empty((world, n), dtype=torch.float32, pin_memory=True) if cuda else recv
                bufs[k] = (recv, host)
            recv, host = bufs[k]
            dist.gather(t, gather_list=list(recv.unbind(0)), dst=0)
            if cuda:
                ev = torch.cuda.Event()
                ev.record()                                  # the current stream has waited for the gather
                with torch.cuda.stream(bufs["_copy_stream"]):
                    bufs["_copy_stream"].wait_event(ev)
                    host.copy_(recv, non_blocking=True)
                    done = torch.cuda.Event()
                    done.record()
                events.append(done)
            out[k] = host
        else:
            dist.gather(t, gather_list=None, dst=0)
    if cuda:
        if rank == 0:
            for ev in events:
                ev.synchronize()
        else:
            torch.cuda.current_stream().synchronize()        # the send has left this rank's OUT block
    return out if rank == 0 else None


class SharedResults:
    """Final-gather buffers in POSIX shared memory: one segment per result array, laid out [world, B_local, ...], so that
    every rank copies its OUT block from its own GPU over its own PCIe link straight into its slice and rank 0 reads
    the concatenated [B_total, ...] arrays without another copy.  Two generations alternate, so the arrays returned by one
    call stay intact while the next call is being written.  All ranks construct it collectively (one node)."""

    def __init__(self, shapes: dict, register: bool = False):
        import torch.distributed as dist
        from multiprocessing import shared_memory

        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        self.shapes, self.gen, self._segs, self.arrays = shapes, 0, [], []
        for g in range(2):
            names = [None] * len(shapes)
            if self.rank == 0:
                segs = [shared_memory.SharedMemory(create=True, size=max(4, self.world * int(np.prod(sh)) * 4)) for sh in shapes.values()]
                names = [sg.name for sg in segs]
            dist.broadcast_object_list(names, src=0)
            if self.rank != 0:
                segs = [shared_memory.SharedMemory(name=nm) for nm in names]
                try:     # attaching registers the segment with this process's resource tracker too (Python < 3.13); rank 0 owns it
                    from multiprocessing import resource_tracker
                    for sg in segs:
                        resource_tracker.unregister(sg._name, "shared_memory")
                except Exception:
                    pass
            self._segs.append(segs)
            self.arrays.append({k: np.ndarray((self.world,) + tuple(sh), np.float32, buffer=sg.buf) for (k, sh), sg in zip(shapes.items(), segs)})
        # page-lock this rank's own slices so that the device-to-host copy goes straight into them (sdempc_fetch_direct);
        # registered ranges are released in close() -- a later mapping may reuse the addresses
        self.registered, self._locked = False, []
        if register:
            from . import _abi
            lib = _abi.load_library()
            ok = True
            for arrs in self.arrays:
                for a in arrs.values():
                    sl = a[self.rank]
                    if lib.sdempc_host_register(sl.ctypes.data, sl.nbytes) == 0:
                        self._locked.append(sl.ctypes.data)
                    else:
                        ok = False
            self.registered = ok
        dist.barrier()
        if self.rank == 0:       # the mappings stay valid; the names disappear at once so nothing leaks if a rank dies
            for segs in self._segs:
                for sg in segs:
                    sg.unlink()

    def next(self) -> dict:
        self.gen ^= 1
        return self.arrays[self.gen]

    def close(self):
        """Release the page locks and the mappings (rank-local; the segments themselves were unlinked at creation)."""
        if self._locked:
            from . import _abi
            lib = _abi.load_library()
            for p in self._locked:
                lib.sdempc_host_unregister(p)
            self._locked = []
        self.registered = False
        self.arrays = []
        for segs in self._segs:
            for sg in segs:
                try:
                    sg.close()
                except Exception:
                    pass
        self._segs = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def solve_sharded(solver, local_problem: dict, u0, info0, B_total: int, gather: str = "shm"):
    """One batched solve of this rank's shard followed by the final result gather to rank 0, the only exchange of the
    path ("host concatenation of per-GPU pinned buffers, or one ncclGather"):

    * ``gather="shm"`` (default): every rank copies its OUT block (x_evol | plan | telemetry) device -> host over its own
      PCIe link into its slice of a shared-memory result array, then one barrier; rank 0 never funnels the other ranks'
      results through its own link (8 x B200: the gather costs what a single GPU's copy costs);
    * ``gather="nccl"``: device-to-device gather over NCCL / NVLink straight from the library's OUT block to rank 0 (one
      gather per sub-array), then rank 0's D2H of all of it, each array's copy overlapped with the next gather.

    Returns the dict of [B_total, ...] arrays on rank 0 (views of reused buffers, valid until the call after next), else None."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [shard_range(B_total, r, world)[1] - shard_range(B_total, r, world)[0] for r in range(world)]
    if len(set(sizes)) != 1:
        raise ValueError("solve_sharded needs equal shards (B_total divisible by the world size)")
    solver.stage(local_problem["x"], u0, info0, xref_win=local_problem.get("xref_win"), rng=local_problem.get("rng"),
                 curr_t=local_problem.get("curr_t"), xdes=local_problem.get("xdes"))
    solver.launch_timed(1, flush_l2=False)          # launch + event wait on the handle's stream
    ptr, nbytes, layout = solver.device_out()
    key = (gather, world, tuple((k, off, shape) for k, (off, shape) in sorted(layout.items())))
    if _gather_cache.get("key") != key:
        if _gather_cache.get("shared") is not None:
            _gather_cache["shared"].close()
        _gather_cache.clear()
        _gather_cache["key"] = key
        _gather_cache["bufs"] = {}
        if gather == "shm":
            _gather_cache["shared"] = SharedResults({k: shape for k, (off, shape) in layout.items()}, register=True)
    if gather == "shm":
        arrs = _gather_cache["shared"].next()
        solver.fetch_into(arrs["u"][rank], arrs["x_evol"][rank], arrs["info"][rank], direct=_gather_cache["shared"].registered)
        dist.barrier()
        if rank != 0:
            return None
        return {k: arrs[k].reshape((world * shape[0],) + tuple(shape[1:])) for k, (off, shape) in layout.items()}
    local = torch.as_tensor(_DevBuf(ptr, nbytes), device="cuda")
    parts = {k: local[off // 4: off // 4 + int(np.prod(shape))] for k, (off, shape) in layout.items()}
    got = gather_to_rank0(parts, _gather_cache["bufs"])   # ~1.5 KB per problem in total
    if got is None:
        return None
    return {k: got[k].numpy().reshape((world * shape[0],) + tuple(shape[1:])) for k, (off, shape) in layout.items()}


def closed_loop_sharded(solver, x0, t0, rng, ticks: int, device=None) -> dict | None:
    """Monte-Carlo closed loop (BASELINE config 5) over all ranks: every rank is given the same global
    ``x0[R,13]``, ``t0[R]``, ``rng[R,2]``, takes its contiguous block of rollouts, runs the whole ``ticks``-tick
    loop of plant + MPC on its GPU in ONE launch (no host synchronisation and no collective inside the loop), and
    the only exchange is at the end: the per-rollout statistics ``[R,4]`` = (rms position error, max position
    error, mean opt_cost, mean num_steps) are gathered on rank 0 and the device time is reduced with MAX.
    Returns on rank 0 ``{"stats", "device_s", "ticks_per_s", "rollouts_per_s", "rollouts_per_rank"}``, else None."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(), dist.get_rank()
    x0 = np.ascontiguousarray(x0, np.float32).reshape(-1, 13)
    R = x0.shape[0]
    lo, hi = shard_range(R, rank, world)
    t0 = np.ascontiguousarray(t0, np.float32).reshape(R)
    rng = np.ascontiguousarray(rng, np.uint64).reshape(R, 2)
    if hi > lo:
        _, _, st = solver.closed_loop(x0[lo:hi], t0[lo:hi], rng[lo:hi], ticks, want_hist=False)
        ms = float(solver.last_launch_ms())
    else:                                    # more ranks than rollouts: this rank only takes part in the exchange
        st, ms = np.zeros((0, 4), np.float32), 0.0
    tms = torch.tensor([ms], dtype=torch.float32)
    if device is not None:
        tms = tms.to(device)
    dist.all_reduce(tms, op=dist.ReduceOp.MAX)       # multi-GPU time = the slowest rank's device time
    res = gather_results({"stats": st}, R, device=device)
    if rank != 0:
        return None
    dev_s = float(tms.item()) * 1e-3
    return {"stats": res["stats"], "device_s": dev_s, "ticks_per_s": R * ticks / dev_s if dev_s > 0 else float("nan"),
            "rollouts_per_s": R / dev_s if dev_s > 0 else float("nan"),
            "rollouts_per_rank": [shard_range(R, r, world)[1] - shard_range(R, r, world)[0] for r in range(world)]}
