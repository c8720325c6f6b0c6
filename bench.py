#!/usr/bin/env python
"""bench.py — ICP iterations/s of the fine-registration hot path (BASELINE.json metric), plus the
36-view chain pairs/s of the same metric string in `extra.chain`.

Workload (configs[1]): point-to-plane ICP (max_corr 0.02 m, 50-iteration cap, PCL convergence
criteria) of a ~307k x ~307k synthetic Kinect-v1 pair 5 degrees apart on the turntable, target
normals from our k=30 normal-estimation pass.  A *step* is one complete pair alignment on
device-resident clouds: spatial-index build over the target + the whole ICP loop +
getFitnessScore.  value = ICP iterations executed / time, summed over steps (index build and
fitness are therefore amortised INTO the number, not excluded).  At N GPUs every rank aligns the
SAME pair (fixed work per GPU = weak scaling of the hot path itself; pairs of different views differ
by +-1 iteration, which would measure load imbalance, not the machine — the multi-pair workload
with its imbalance is the chain arm below); nothing crosses NVLink inside a step — the per-step
4x4 records stay on the rank and are gathered ONCE after the timed region (one NCCL all_gather),
rank 0 checks that all GPUs produced bit-identical poses: value = all ranks' iterations / max time.

`extra.chain` (configs[2]): the 36-view turntable chain — per view VoxelGrid 2 mm + SOR k=50 +
normals k=30 chained on the device (lc3d_prepare_view), per pair point-to-plane ICP on the
resident views — sharded over the N ranks in contiguous pair blocks, one gather per chain:
pairs/s = 35 x repetitions / max-over-ranks wall time (strong scaling: the chain is fixed).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

STEP_DEG = 5.0
MAX_CORR = 0.02
MAX_ITER = 50
K_NORMALS = 30
WORKLOAD = ("point-to-plane ICP, 640x480 synthetic Kinect-v1 full-frame pair (~307k x ~307k pts, 5 deg "
            "turntable step), max_corr=0.02 m, k=30 normals, 50-iteration cap")
CHAIN_VIEWS = 36
CHAIN_STEP = 360.0 / CHAIN_VIEWS
CHAIN_LEAF, CHAIN_SOR_K, CHAIN_SOR_MUL = 0.002, 50, 1.0
CHAIN_PREFETCH = int(os.environ.get("LC3D_CHAIN_PREFETCH", "1"))  # views prepared ahead on a worker thread (0 = serial)
CHAIN_LANES = int(os.environ.get("LC3D_CHAIN_LANES", "3"))  # host threads (pair sub-blocks) per GPU
CHAIN_MODE = os.environ.get("LC3D_CHAIN_MODE", "native")  # native | dag | lanes
CHAIN_PREP = int(os.environ.get("LC3D_CHAIN_PREP", "4"))    # dag mode: view-preparation threads per GPU
CHAIN_ALIGN = int(os.environ.get("LC3D_CHAIN_ALIGN", "3"))  # dag mode: pair-alignment threads per GPU


# ------------------------------------------------------------------------------------ inputs

def _render(job):
    from lowcost3dreconstruction_b200 import synth
    view, step, backdrop = job
    return synth.kinect_view(view, step_deg=step, backdrop=backdrop)


def render_views(jobs):
    """[(view, step_deg, backdrop)] -> clouds; cached under /tmp, rendered in a process pool
    (numpy ray marching, ~1-2 s per view on one core).  Must run BEFORE CUDA is initialised."""
    out, todo = {}, []
    for j in jobs:
        path = f"/tmp/lc3d_bench_view_{j[0]}_{j[1]}_{j[2]}.npy"
        if os.path.exists(path):
            out[j] = np.load(path)
        else:
            todo.append(j)
    if todo:
        world = int(os.environ.get("WORLD_SIZE", "1"))
        procs = max(1, min(len(todo), (os.cpu_count() or 1) // max(world, 1), 16))
        if procs > 1:
            with mp.get_context("fork").Pool(procs) as pool:
                res = pool.map(_render, todo)
        else:
            res = [_render(j) for j in todo]
        for j, c in zip(todo, res):
            out[j] = c
            try:
                np.save(f"/tmp/lc3d_bench_view_{j[0]}_{j[1]}_{j[2]}.npy", c)
            except OSError:
                pass
    return out


def load_pair(view: int):
    """(source = view+1, target = view) clouds of the bench pair."""
    v = render_views([(view + 1, STEP_DEG, "full"), (view, STEP_DEG, "full")])
    return v[(view + 1, STEP_DEG, "full")], v[(view, STEP_DEG, "full")]


def host_info() -> dict:
    model = "unknown"
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    model = line.split(":", 1)[1].strip()
                    break
    except OSError:
        pass
    return {"host_cores": os.cpu_count(), "cpu_model": model}


class ClockSampler:
    """SM clock / throttle reasons polled through NVML (every ~2 ms) while the timed region runs."""

    def __init__(self, gpu: int):
        self.gpu, self.rows, self.stop_flag, self.thread, self.h = gpu, [], False, None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(gpu)
        except Exception as e:  # noqa: BLE001
            self.err = str(e)

    def _poll(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                self.rows.append((nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM),
                                  nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)))
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.002)

    def start(self):
        if self.h is not None:
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()

    def stop(self) -> dict:
        if self.h is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"], "samples": 0}
        self.stop_flag = True
        self.thread.join(timeout=1.0)
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        reasons = sorted({nm for _, bits in self.rows for nm, m in names.items() if bits & m})
        sm = [c for c, _ in self.rows]
        try:
            smax = float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM))
        except Exception:  # noqa: BLE001
            smax = None
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": reasons,
                "samples": len(sm)}


# ----------------------------------------------------------------------------- reference arm

def run_reference(args, rank: int, world: int):
    """--impl reference: the reference's CPU path for this workload.  PCL cannot be built in
    this image (no PCL/Boost/Eigen/FLANN), so this times the oracle — the CPU restatement of
    the PCL algorithm (kind "port") — single-threaded like pcl::IterativeClosestPoint.
    A step is the same unit as in the GPU arm: ONE complete alignment of the same pair
    (kd-tree build over the target + the whole ICP loop + getFitnessScore); value = ICP
    iterations executed / time.  (~1.2 s per step.)"""
    if rank != 0:
        return
    from lowcost3dreconstruction_b200._capi import HostCloud
    from oracle import oracle as orc
    src, tgt = load_pair(0)
    if os.path.exists("/tmp/lc3d_bench_nrm0.npy"):
        nrm = np.load("/tmp/lc3d_bench_nrm0.npy")
    else:
        nrm, _ = orc.normals(tgt, K_NORMALS)
    T = HostCloud(tgt, normal=nrm)
    S = HostCloud(src)
    for _ in range(min(args.warmup, 1)):
        orc.icp_align(S, T, MAX_CORR, MAX_ITER, mode=1, compute_fitness=True)
    iters = 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        iters += orc.icp_align(S, T, MAX_CORR, MAX_ITER, mode=1, compute_fitness=True)["iterations"]
    dt = time.perf_counter() - t0
    val = iters / dt
    line = {
        "impl": "reference", "metric": "icp_iters_per_sec", "value": val, "unit": "iterations/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "n_source": int(S.n), "n_target": int(T.n), "mode": "point-to-plane",
                   "step": "kd-tree build + ICP loop + fitness on host arrays"},
        "cpu_baseline": dict({"value": val, "unit": "iterations/s", "cores": 1, "kind": "port",
                              "sample": f"{args.steps} complete alignments of the same pair (kd-tree build + "
                                        f"{iters // max(args.steps, 1)} iterations + fitness each); oracle restatement "
                                        "of PCL (PCL itself cannot be built here), 1 thread like "
                                        "pcl::IterativeClosestPoint"}, **host_info()),
        "e2e": {"value": val, "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# -------------------------------------------------------------------------------- chain arm

def chain_inputs(rank: int, world: int):
    """Host-side inputs of this rank's block of the 36-view chain: raw clouds pre-aligned with the
    turntable prior of rotate_align (rotate_align.cpp:224-235) perturbed by a residual error."""
    from lowcost3dreconstruction_b200 import chain, synth
    pairs = chain.shard_pairs(CHAIN_VIEWS - 1, world, rank)
    needed = sorted({p for p in pairs} | {p - 1 for p in pairs})
    rng = np.random.default_rng(7)
    resid = {v: synth.rigid(*(rng.normal(0, 0.4, 3)), rng.normal(0, 0.002, 3)) for v in range(CHAIN_VIEWS)}
    clouds = render_views([(v, CHAIN_STEP, "none") for v in needed])
    raw = {v: synth.apply_transform(resid[v] @ synth.turntable_prior(v, CHAIN_STEP), clouds[(v, CHAIN_STEP, "none")])
           for v in needed}
    return pairs, raw, resid


def run_chain(ctx, rank, world, dev, pairs, raw, resid, reps: int) -> dict | None:
    import torch
    import torch.distributed as dist
    from lowcost3dreconstruction_b200 import api, chain

    # The rank's block runs as a task graph on its GPU (chain.align_pairs_dag): CHAIN_PREP host threads
    # prepare views, CHAIN_ALIGN threads align pairs as soon as both views are ready; every thread has
    # its own lc3d context (own stream, own scratch).  LC3D_CHAIN_MODE=lanes selects the sub-block
    # pipelines of chain.align_pairs_lanes instead.
    native = CHAIN_MODE == "native"  # the same task graph run by the library's own host threads (lc3d_chain_run)
    dag = CHAIN_MODE == "dag" or native
    n_lanes = max(1, min(CHAIN_LANES, len(pairs) // 2))
    n_prep = max(1, min(CHAIN_PREP, len(pairs) + 1)) if dag else n_lanes
    n_align = max(1, min(CHAIN_ALIGN, len(pairs))) if dag else n_lanes
    icp_ctx = [ctx] + ([] if native else [api.Context(ctx.device) for _ in range(n_align - 1)])
    prep_ctx = [] if native else [api.Context(ctx.device) for _ in range(n_prep)]
    pinned = {v: api.host_register(a) for v, a in raw.items()}  # the PLY loader's buffers, pinned once
    nat = chain.NativeChain(ctx.device, n_prep, n_align) if native else None
    nat_views = [pinned[v] for v in range(pairs[0] - 1, pairs[-1] + 1)] if native and pairs else []
    nat_kw = dict(leaf_size=CHAIN_LEAF, sor_mean_k=CHAIN_SOR_K, sor_stddev_mul=CHAIN_SOR_MUL, normals_k=K_NORMALS,
                  max_correspondence_distance=MAX_CORR, max_iterations=MAX_ITER, mode=api.POINT_TO_PLANE)

    def one_chain():
        """this rank's views prepared on the device, its pairs aligned resident"""
        local, npts = np.zeros((CHAIN_VIEWS - 1, chain.RECORD)), []

        def prep_fn(i):
            def get_view(v):
                d, cnt = api.prepare_view(pinned[v], CHAIN_LEAF, CHAIN_SOR_K, CHAIN_SOR_MUL, K_NORMALS, ctx=prep_ctx[i])
                npts.append(cnt[2])
                return d
            return get_view

        def align_fn(i):
            def align(src, tgt):
                return api.icp_align(src, tgt, MAX_CORR, MAX_ITER, mode=api.POINT_TO_PLANE, ctx=icp_ctx[i])
            return align

        if native:
            if nat_views:
                res_, np_ = nat.run(nat_views, **nat_kw)
                for i_, r_ in enumerate(res_):
                    local[pairs[0] - 1 + i_] = chain.pack_record(r_)
                npts.extend(np_)
        elif dag:
            chain.align_pairs_dag(pairs, [prep_fn(i) for i in range(n_prep)], [align_fn(i) for i in range(n_align)],
                                  local, release=lambda d: d.free())
        else:
            chain.align_pairs_lanes(pairs, [(prep_fn(i), align_fn(i), (lambda d: d.free())) for i in range(n_lanes)],
                                    local, prefetch=CHAIN_PREFETCH)
        return chain.exchange_records(local, dev), npts  # one gather per chain (20 doubles per pair)

    # the lanes are Python threads whose work is inside GIL-releasing C calls; a short switch interval
    # keeps the hand-offs between them from waiting on the interpreter's default 5 ms tick
    switch0 = sys.getswitchinterval()
    sys.setswitchinterval(5e-5)
    # warm-up.  Which context handles which view / pair is decided at run time, and the scratch buffers
    # of a context grow to the largest cloud / cell table it has seen (a regrowth is a cudaFree: a
    # device-wide synchronisation, 5-25 ms when seven streams are busy): every context sees every view /
    # pair of the block once, so that the timed repetitions allocate nothing
    # (rank0_device_allocations_per_repetition).  Then two whole chains (NCCL communicator, pools).
    if native:
        if nat_views:
            nat.run(nat_views, warm=True, **nat_kw)
    elif dag:
        for c in prep_ctx:
            held = {v: api.prepare_view(pinned[v], CHAIN_LEAF, CHAIN_SOR_K, CHAIN_SOR_MUL, K_NORMALS, ctx=c)[0]
                    for v in sorted({p for p in pairs} | {p - 1 for p in pairs})}
            if c is prep_ctx[-1]:
                for ic in icp_ctx:
                    for p in pairs:
                        api.icp_align(held[p], held[p - 1], MAX_CORR, MAX_ITER, mode=api.POINT_TO_PLANE, ctx=ic)
            for d in held.values():
                d.free()
    for _ in range(2):
        one_chain()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    rep_s, rep_allocs = [], []
    for _ in range(reps):
        r0, a0 = time.perf_counter(), ctx._lib.lc3d_debug_alloc_count()
        rec, npts = one_chain()
        rep_s.append(time.perf_counter() - r0)
        rep_allocs.append(int(ctx._lib.lc3d_debug_alloc_count() - a0))
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    sys.setswitchinterval(switch0)
    tt = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    for a in pinned.values():
        api.host_unregister(a)
    for c in prep_ctx + icp_ctx[1:]:
        c.close()
    if nat:
        nat.close()
    if rank != 0:
        return None
    dt = float(tt[0])
    res = [chain.unpack_record(r) for r in rec]
    poses = chain.compose_chain([p["transformation"] for p in res])
    errs = []
    for p, pr in enumerate(res, start=1):  # ground truth: pair p's residual motion is resid[p-1] . resid[p]^-1
        Tt = resid[p - 1] @ np.linalg.inv(resid[p])
        R = pr["transformation"][:3, :3].astype(np.float64) @ Tt[:3, :3].T
        errs.append(float(np.degrees(np.arccos(np.clip((np.trace(R) - 1) / 2, -1, 1)))))
    return {"pairs_per_sec": (CHAIN_VIEWS - 1) * reps / dt, "unit": "pairs/s", "scaling": "strong", "views": CHAIN_VIEWS,
            "pairs": CHAIN_VIEWS - 1, "repetitions": reps, "seconds_per_chain": dt / reps,
            "iterations_total": int(sum(p["iterations"] for p in res)),
            "converged_pairs": int(sum(p["converged"] for p in res)),
            "median_rot_err_deg_vs_truth": float(np.median(errs)), "max_rot_err_deg_vs_truth": float(np.max(errs)),
            "points_per_view_after_voxel_sor": int(np.mean(npts)) if npts else 0,
            "pose_35_translation_m": [float(x) for x in poses[-1][:3, 3]],
            "schedule": (f"task graph: {n_prep} view-preparation + {n_align} pair-alignment host threads per GPU"
                         + (" (native executor, lc3d_chain_run)" if native else " (Python threads)") if dag
                         else f"{n_lanes} sub-block pipelines per GPU, prefetch {CHAIN_PREFETCH}"),
            "prefetch_views": CHAIN_PREFETCH, "lanes_per_gpu": n_lanes,
            "rank0_seconds_per_repetition": [round(x, 5) for x in rep_s],
            "rank0_device_allocations_per_repetition": rep_allocs,
            "timed": "per view VoxelGrid 2 mm + SOR k=50 + normals k=30 on the device (lc3d_prepare_view, page-locked host "
                     "xyz in, on a second context one view ahead of the pair being aligned), per pair point-to-plane "
                     "ICP on the resident views, one record gather per chain; wall clock bracketed by barrier + cuda "
                     "synchronize, max over ranks"}


# ---------------------------------------------------------------------------------- main arm

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-chain", action="store_true")
    ap.add_argument("--chain-reps", type=int, default=5)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    # ---- inputs are rendered before CUDA comes up (fork-based process pool) --------------------
    src, tgt = load_pair(0)  # the same pair on every rank (see the module docstring)
    chain_in = None if args.no_chain else chain_inputs(rank, world)

    import torch
    import torch.distributed as dist
    from lowcost3dreconstruction_b200 import api
    from lowcost3dreconstruction_b200._capi import HostCloud

    # host threads and pinned buffers next to the GPU: NVML's CPU affinity of the device (the cores of
    # its NUMA node).  The e2e arm is a host-side wall clock over PCIe copies from pinned memory; a
    # process that lands on the far socket sees 30 % less of it.
    affinity, gpu_uuid = None, None
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        gpu_uuid = pynvml.nvmlDeviceGetUUID(h)
        gpu_uuid = gpu_uuid.decode() if isinstance(gpu_uuid, bytes) else str(gpu_uuid)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cores = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1}
        cores &= set(os.sched_getaffinity(0))
        if cores:
            os.sched_setaffinity(0, cores)
            affinity = len(cores)
    except Exception:  # noqa: BLE001 - best effort (no NVML, containers without the syscall)
        pass
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    # one explicit stream for everything that is timed: our kernels, the L2 flush and the events
    stream = torch.cuda.Stream(device=dev)
    ctx = api.Context(local_rank, stream=stream.cuda_stream)

    # ---- the bench pair: view 1 -> view 0 ---------------------------------------------------
    n_t, c_t = api.normals(tgt, K_NORMALS, ctx=ctx)
    n_s, c_s = api.normals(src, K_NORMALS, ctx=ctx)
    t0 = time.perf_counter()
    api.normals(tgt, K_NORMALS, ctx=ctx)
    ms_normals = (time.perf_counter() - t0) * 1e3
    if rank == 0:
        try:
            np.save("/tmp/lc3d_bench_nrm0.npy", n_t)
        except OSError:
            pass
    S = HostCloud(src, normal=n_s, curvature=c_s)
    T = HostCloud(tgt, normal=n_t, curvature=c_t)
    dS, dT = ctx.upload(S), ctx.upload(T)
    with torch.cuda.stream(stream):
        flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def one_step():
        return api.icp_align(dS, dT, MAX_CORR, MAX_ITER, mode=api.POINT_TO_PLANE, compute_fitness=True, ctx=ctx)

    sampler = ClockSampler(local_rank)
    sampler.start()  # covers warm-up + timed region (the timed region alone lasts ~20 ms)
    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            flush.zero_()
            one_step()
        # ---- timed region: exactly K steps, device-timed, L2 flushed between steps ----------
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        l0 = ctx.launch_count
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        results = []
        wall0 = time.perf_counter()
        for a, b in ev:
            flush.zero_()
            a.record(stream)
            results.append(one_step())
            b.record(stream)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        wall = time.perf_counter() - wall0
    launches = ctx.launch_count - l0
    # every step aligns the same resident pair: anything but identical bits would be a defect
    steps_identical = all(np.array_equal(r["transformation"], results[0]["transformation"]) and
                          r["fitness"] == results[0]["fitness"] and r["iterations"] == results[0]["iterations"] and
                          r["last_correspondences"] == results[0]["last_correspondences"] for r in results)
    clocks = sampler.stop()
    ms_steps = sum(a.elapsed_time(b) for a, b in ev)
    iters = sum(r["iterations"] for r in results)
    ms_loop = sum(r["ms"]["loop"] for r in results)
    ms_index = sum(r["ms"]["index"] for r in results)
    ms_fit = sum(r["ms"]["fitness"] for r in results)
    t = torch.tensor([ms_steps, float(iters), ms_loop], dtype=torch.float64, device=dev)
    tmax, tsum = t.clone(), t.clone()
    gather_ms = 0.0
    ranks_identical = True
    per_rank_ms = [ms_steps / args.steps]
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        allt = torch.empty(world, 3, dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(allt.view(-1), t)
        per_rank_ms = [float(x) / args.steps for x in allt[:, 0].cpu()]
        # the one exchange of the run: every step's 4x4 (+ fitness, iterations) from every rank;
        # rank 0 checks that every GPU produced the same bits
        mine = torch.tensor(np.stack([np.concatenate([r["transformation"].reshape(16).astype(np.float64),
                                                      [r["fitness"], r["iterations"], 0.0, 0.0]]) for r in results]),
                            device=dev)
        allrec = torch.empty(world, *mine.shape, dtype=mine.dtype, device=dev)
        torch.cuda.synchronize()
        g0 = time.perf_counter()
        dist.all_gather_into_tensor(allrec.view(-1), mine.view(-1))
        torch.cuda.synchronize()
        gather_ms = (time.perf_counter() - g0) * 1e3
        if rank == 0:
            rec = allrec.cpu().numpy()
            ranks_identical = bool(all(np.array_equal(rec[r], rec[0]) for r in range(world)))
    ms_max = float(tmax[0])
    total_iters = float(tsum[1])
    value = total_iters / (ms_max * 1e-3)

    # ---- e2e: the reference-facing host-buffer C-ABI call --------------------------------------
    def reduce_e2e(e_t, e_iters):
        e = torch.tensor([e_t, float(e_iters)], dtype=torch.float64, device=dev)
        emax, esum = e.clone(), e.clone()
        if world > 1:
            dist.all_reduce(emax, op=dist.ReduceOp.MAX)
            dist.all_reduce(esum, op=dist.ReduceOp.SUM)
        return float(esum[1]) / float(emax[0])

    def time_e2e(Sx, Tx, reg_out, reps):
        for _ in range(2):
            api.icp_align(Sx, Tx, MAX_CORR, MAX_ITER, mode=api.POINT_TO_PLANE, registered_out=reg_out, ctx=ctx)
        e_iters, e_t = 0, 0.0
        for _ in range(reps):
            with torch.cuda.stream(stream):
                flush.zero_()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            r = api.icp_align(Sx, Tx, MAX_CORR, MAX_ITER, mode=api.POINT_TO_PLANE, registered_out=reg_out, ctx=ctx)
            dt_ = time.perf_counter() - t0
            e_t += dt_
            e2e_step_ms.append(dt_ * 1e3)
            e_iters += r["iterations"]
        return reduce_e2e(e_t, e_iters)

    def pinned(a):
        t_ = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t_, t_.numpy()
    chain_res = None
    if chain_in is not None and os.environ.get("LC3D_BENCH_CHAIN_FIRST"):
        chain_res = run_chain(ctx, rank, world, dev, *chain_in, reps=max(1, args.chain_reps))
        chain_in = None
    e_reps = max(3, min(args.steps, 20))
    e2e_step_ms: list[float] = []

    def pcie_rates():
        a = torch.empty(64 << 20, dtype=torch.uint8).pin_memory()
        d = torch.empty(64 << 20, dtype=torch.uint8, device=dev)
        out = []
        for src_, dst_ in ((a, d), (d, a)):
            for _ in range(2):
                dst_.copy_(src_, non_blocking=True)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(4):
                dst_.copy_(src_, non_blocking=True)
            torch.cuda.synchronize()
            out.append(4 * 64 / 1024 / (time.perf_counter() - t0))
        return out
    pcie_h2d, pcie_d2h = pcie_rates()
    # (a) headline: packed arrays in pinned host memory (what a caller that owns its buffers does)
    keep = [pinned(x) for x in (src, n_s, tgt, n_t, np.empty_like(src), np.empty_like(src))]
    Sp = HostCloud(keep[0][1], normal=keep[1][1])
    Tp = HostCloud(keep[2][1], normal=keep[3][1])
    e2e_val = time_e2e(Sp, Tp, (keep[4][1], keep[5][1]), e_reps)
    e2e_median_ms = float(np.median(e2e_step_ms))  # (host wall clock: the median shows the jitter of the mean)
    h2d = int(Sp.n * 24 + Tp.n * 24)
    d2h = int(Sp.n * 24 + 128)
    # (b) what INTEGRATION.md's in-main() binding passes: PCL's own 48-byte PointXYZRGBNormal array in
    # pageable memory (std::vector), registered cloud back into pageable memory
    def aos(xyz, nrm, curv):
        a = np.zeros((len(xyz), 12), dtype=np.float32)
        a[:, 0:3], a[:, 3], a[:, 4:7], a[:, 9] = xyz, 1.0, nrm, curv
        return a
    Sa, Ta = HostCloud.from_pcl_aos(aos(src, n_s, c_s)), HostCloud.from_pcl_aos(aos(tgt, n_t, c_t))
    e2e_aos = time_e2e(Sa, Ta, (np.empty_like(src), np.empty_like(src)), e_reps)

    # ---- 36-view chain (configs[2]) -----------------------------------------------------------
    if chain_in is not None:
        chain_res = run_chain(ctx, rank, world, dev, *chain_in, reps=max(1, args.chain_reps))

    # ---- configs[0]: the reference's own mode — point-to-point, ~200k-point pair, 50 iterations ----
    cfg1 = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import oracle as orc
        v1 = render_views([(1, STEP_DEG, "panel"), (0, STEP_DEG, "panel")])
        s1, t1 = v1[(1, STEP_DEG, "panel")], v1[(0, STEP_DEG, "panel")]
        d1s, d1t = ctx.upload(HostCloud(s1)), ctx.upload(HostCloud(t1))
        for _ in range(2):
            g1 = api.icp_align(d1s, d1t, MAX_CORR, MAX_ITER, mode=api.POINT_TO_POINT, ctx=ctx)
        reps1, ms1 = 5, 0.0
        for _ in range(reps1):
            g1 = api.icp_align(d1s, d1t, MAX_CORR, MAX_ITER, mode=api.POINT_TO_POINT, ctx=ctx)
            ms1 += g1["ms"]["total"]
        t0 = time.perf_counter()
        o1 = orc.icp_align(s1, t1, MAX_CORR, MAX_ITER, mode=0)
        dt1 = time.perf_counter() - t0
        o1f = orc.icp_align(s1, t1, MAX_CORR, MAX_ITER, mode=0, umeyama_f32=True)  # PCL's float32 Umeyama sums

        def delta(o):
            return {"iterations": int(o["iterations"]), "state": int(o["state"]),
                    "transform_max_abs_diff_vs_gpu": float(np.abs(o["transformation"] - g1["transformation"]).max()),
                    "fitness_rel_diff_vs_gpu": float(abs(o["fitness"] - g1["fitness"]) / o["fitness"])}
        cfg1 = {"workload": "point-to-point ICP (pcl_tools/fine_registration.cpp:105), synthetic Kinect-v1 pair with "
                            "finite backdrop, 5 deg step, max_corr 0.02 m, 50-iteration cap",
                "n_source": int(len(s1)), "n_target": int(len(t1)), "gpu_iterations": int(g1["iterations"]),
                "gpu_state": int(g1["state"]), "gpu_iters_per_sec_resident": g1["iterations"] * reps1 / (ms1 * 1e-3),
                "gpu_ms_per_alignment": ms1 / reps1, "cpu_iters_per_sec": o1["iterations"] / dt1,
                "oracle_fp64_sums": delta(o1), "oracle_float32_sums_like_pcl": delta(o1f)}
        d1s.free()
        d1t.free()

    # ---- roofline of the dominant kernel (the fused ICP iteration) ---------------------------
    peaks, prof = {}, {}
    for name, dst in (("MEASURED_PEAKS.json", peaks), (os.path.join("profiles", "r02_roofline_traffic.json"), prof)):
        try:
            with open(os.path.join(ROOT, name)) as f:
                dst.update(json.load(f))
        except (OSError, ValueError):
            pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    t_iter_s = (ms_loop / max(iters, 1)) * 1e-3
    alg_bytes = 64.0 * S.n  # whole point-to-plane iteration = 64 B per source point
    achieved = alg_bytes / t_iter_s / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                # dram__bytes_read.sum + dram__bytes_write.sum per launch of a converged iteration, from the
                # committed `ncu --set full` capture named in that file (caches flushed per replay)
                "traffic": prof.get("dram_bytes_per_launch"), "traffic_source": prof.get("source"),
                "kernel": "icp_iteration_kernel<point-to-plane> (+ icp_solve_kernel)",
                "algorithmic_bytes_per_launch": alg_bytes, "us_per_launch": t_iter_s * 1e6,
                "peak_source": "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback 6.65 TB/s"}

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            from oracle import oracle as orc
            t0 = time.perf_counter()
            o = orc.icp_align(S, T, MAX_CORR, MAX_ITER, mode=1, compute_fitness=True)
            dt = time.perf_counter() - t0
            g = results[-1]
            cpu = dict({"value": o["iterations"] / dt, "unit": "iterations/s", "cores": 1, "kind": "port",
                        "sample": f"one full alignment of the same pair (kd-tree build + {o['iterations']} iterations "
                                  f"+ fitness) = {dt:.1f} s; oracle restatement of PCL, 1 thread",
                        "iterations": o["iterations"],
                        "transform_max_abs_diff_vs_gpu": float(np.abs(o["transformation"] - g["transformation"]).max()),
                        "fitness_rel_diff_vs_gpu": float(abs(o["fitness"] - g["fitness"]) / o["fitness"])},
                       **host_info())
            cpu["cfg1_point_to_point"] = cfg1
        line = {
            "metric": "icp_iters_per_sec", "value": value, "unit": "iterations/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "n_source": int(S.n), "n_target": int(T.n), "mode": "point-to-plane",
                       "pairs_per_gpu": 1, "pair": "the same pair (views 1 -> 0) on every GPU", "l2": "flushed between steps (256 MiB write on the timed stream)",
                       "step": "index build + ICP loop + fitness on resident clouds"},
            "roofline": roofline, "cpu_baseline": cpu,
            "e2e": {"value": e2e_val, "unit": "iterations/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "host_memory": "pinned, packed xyz / normal arrays",
                    "pcie_pinned_gbs": {"h2d": round(pcie_h2d, 1), "d2h": round(pcie_d2h, 1)},
                    "cpu_affinity_cores": affinity, "steps": e_reps, "median_ms_per_step": round(e2e_median_ms, 4)},
            "gpu_launches": int(launches), "clocks": clocks,
            "extra": {"iterations_per_alignment": iters / args.steps, "pairs_per_sec": world * args.steps / (ms_max * 1e-3),
                      "ms_index_per_step": ms_index / args.steps, "ms_loop_per_step": ms_loop / args.steps,
                      "ms_fitness_per_step": ms_fit / args.steps, "loop_only_iters_per_sec": iters / (ms_loop * 1e-3),
                      "ms_normals_k30_host_call": ms_normals, "wall_s_timed_region": wall,
                      "fitness": results[-1]["fitness"], "state": results[-1]["state"],
                      "record_gather_ms_after_timed_region": gather_ms,
                      "all_ranks_bit_identical_results": ranks_identical, "all_steps_bit_identical_results": bool(steps_identical),
                      "last_correspondences": int(results[-1]["last_correspondences"]), "gpu_uuid_rank0": gpu_uuid,
                      "ms_per_step_by_rank": [round(x, 4) for x in per_rank_ms],
                      "e2e_pcl_aos_pageable": {"value": e2e_aos, "unit": "iterations/s",
                                               "h2d_bytes_per_step": int(Sa.n * 48 + Ta.n * 48), "d2h_bytes_per_step": d2h,
                                               "host_memory": "pageable, 48-byte pcl::PointXYZRGBNormal AoS "
                                                              "(INTEGRATION.md section B binding)"},
                      "chain": chain_res},
        }
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
