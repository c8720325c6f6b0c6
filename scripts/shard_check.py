#!/usr/bin/env python
"""Source-sharded single-pair ICP (SURVEY 8e second mode) checked against the one-GPU run, and timed.

  torchrun --nproc-per-node N scripts/shard_check.py [points]        (N GPUs, or N processes on one GPU
                                                                      with LC3D_SHARD_ONE_GPU=1)
Every rank builds the same seeded pair, holds the whole target and its slice of the source.  Checked:
iteration count / convergence state equal to the unsharded run, correspondences of iteration 0 and 2
bit-identical to its slices, pose within 1e-6, fitness within 1e-9 relative."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import torch.distributed as dist
    from lowcost3dreconstruction_b200 import api, chain, synth
    from lowcost3dreconstruction_b200._capi import HostCloud

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = 0 if os.environ.get("LC3D_SHARD_ONE_GPU") else int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("gloo" if os.environ.get("LC3D_SHARD_ONE_GPU") else "nccl",
                            **({} if os.environ.get("LC3D_SHARD_ONE_GPU") else {"device_id": torch.device("cuda", local)}))
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 2_000_000
    tgt = synth.surface_samples(n, seed=5)
    T = synth.rigid(0.0, 3.0, 0.0, [0.005, 0.0, 0.0])
    src = synth.apply_transform(np.linalg.inv(T), synth.surface_samples(n, seed=6))
    ctx = api.Context(local)
    nrm, curv = api.normals(tgt, 30, ctx=ctx)
    dT = ctx.upload(HostCloud(tgt, normal=nrm, curvature=curv))
    sp = chain.ShardedPair(ctx, max_shard_points=n // world + 1)
    sl = sp.shard(n)
    dS = ctx.upload(HostCloud(src[sl]))
    res, ok = {}, True
    for mode in (1, 0):
        iters = 12
        for it in (0, 2):
            r = sp.align(dS, dT, 0.02, iters, mode=mode, dump_iteration=it)
            res[(mode, it)] = r
        # timing: 3 alignments
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            r = sp.align(dS, dT, 0.02, iters, mode=mode)
        dt = (time.perf_counter() - t0) / 3
        res[(mode, "time")] = (dt, r)
    # the unsharded run on rank 0
    if rank == 0:
        dF = ctx.upload(HostCloud(src))
        out = {"points": n, "world": world}
        for mode in (1, 0):
            iters = 12
            for it in (0, 2):
                f = api.icp_align(dF, dT, 0.02, iters, mode=mode, dump_iteration=it, ctx=ctx)
                s = res[(mode, it)]
                ok &= np.array_equal(f["corr_index"][sl], s["corr_index"])
                ok &= (f["iterations"], f["state"]) == (s["iterations"], s["state"])
                ok &= float(np.abs(f["transformation"] - s["transformation"]).max()) < 1e-6
                ok &= abs(f["fitness"] - s["fitness"]) <= 1e-9 * f["fitness"]
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(3):
                f = api.icp_align(dF, dT, 0.02, iters, mode=mode, ctx=ctx)
            dtf = (time.perf_counter() - t0) / 3
            dts, s = res[(mode, "time")]
            out[f"mode{mode}"] = {"iterations": s["iterations"], "ms_sharded_wall": dts * 1e3, "ms_one_gpu_wall": dtf * 1e3,
                                  "ms_loop_sharded": s["ms"]["loop"], "ms_loop_one_gpu": f["ms"]["loop"],
                                  "ms_index_sharded": s["ms"]["index"], "ms_index_one_gpu": f["ms"]["index"],
                                  "us_per_iteration_sharded": s["ms"]["loop"] / max(s["iterations"], 1) * 1e3,
                                  "us_per_iteration_one_gpu": f["ms"]["loop"] / max(f["iterations"], 1) * 1e3,
                                  "max_abs_T_diff": float(np.abs(f["transformation"] - s["transformation"]).max())}
        out["parity_ok"] = bool(ok)
        out["exchange_bytes_per_iteration_per_peer"] = "29 (17) rows of <= nblk doubles read over NVLink by 29 (17) blocks"
        print(json.dumps(out), flush=True)
    dist.barrier()
    sp.close()
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
