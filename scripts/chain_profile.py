"""Stage breakdown of the 36-view chain on one GPU (developer tool): wall time of prepare_view
(+ LC3D_PREP_TIMING stage lines on stderr) and of the resident pair alignments."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from lowcost3dreconstruction_b200 import api, chain
pairs, raw, resid = bench.chain_inputs(0, 1) if hasattr(bench, "chain_inputs") else None
ctx = api.Context(0)
for rep in range(3):
    tp = ti = 0.0
    prepared, iters, ms = {}, 0, np.zeros(6)
    t00 = time.perf_counter()
    for p in pairs:
        for v in (p, p - 1):
            if v not in prepared:
                t0 = time.perf_counter()
                prepared[v], cnt = api.prepare_view(raw[v], bench.CHAIN_LEAF, bench.CHAIN_SOR_K, bench.CHAIN_SOR_MUL, bench.K_NORMALS, ctx=ctx)
                tp += time.perf_counter() - t0
        t0 = time.perf_counter()
        r = api.icp_align(prepared[p], prepared[p - 1], bench.MAX_CORR, bench.MAX_ITER, mode=api.POINT_TO_PLANE, ctx=ctx)
        ti += time.perf_counter() - t0
        iters += r["iterations"]
        ms += np.array([r["ms"][k] for k in ("upload", "index", "loop", "fitness", "download", "total")])
        prepared.pop(p - 1).free()
    for d in prepared.values():
        d.free()
    tot = time.perf_counter() - t00
    print(f"rep {rep}: chain {tot*1e3:.1f} ms | prepare_view {tp*1e3:.1f} ms ({tp/36*1e3:.2f}/view) | icp_align {ti*1e3:.1f} ms "
          f"({ti/35*1e3:.2f}/pair, {iters} iterations) | device ms/pair: index {ms[1]/35:.3f} loop {ms[2]/35:.3f} fitness {ms[3]/35:.3f} total {ms[5]/35:.3f}",
          flush=True)
