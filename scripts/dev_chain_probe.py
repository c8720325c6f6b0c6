"""Developer probe: where a chain view's time goes (prepare stages vs ICP)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
pairs, raw, resid = bench.chain_inputs(0, 6)
import torch
from lowcost3dreconstruction_b200 import api
ctx = api.Context(0)
vs = sorted(raw)
def t(f, n=3):
    best = 1e9
    for _ in range(n):
        torch.cuda.synchronize(); t0 = time.perf_counter(); r = f(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    return best * 1e3, r
c = raw[vs[0]]
print("points", len(c))
print("prepare all   ms", t(lambda: api.prepare_view(c, 0.002, 50, 1.0, 30, ctx=ctx))[0])
print("voxel only    ms", t(lambda: api.prepare_view(c, 0.002, 0, 1.0, 0, ctx=ctx))[0])
ms, (d, cnt) = t(lambda: api.prepare_view(c, 0.002, 0, 1.0, 0, ctx=ctx)); vox = d.download()[0]
print("sor only      ms", t(lambda: api.prepare_view(vox, 0.0, 50, 1.0, 0, ctx=ctx))[0], len(vox))
print("normals only  ms", t(lambda: api.prepare_view(vox, 0.0, 0, 1.0, 30, ctx=ctx))[0])
print("host voxel    ms", t(lambda: api.voxel_grid(c, 0.002, ctx=ctx))[0])
print("host sor      ms", t(lambda: api.sor(vox, 50, 1.0, ctx=ctx))[0])
print("host normals  ms", t(lambda: api.normals(vox, 30, ctx=ctx))[0])
a, _ = api.prepare_view(raw[vs[1]], 0.002, 50, 1.0, 30, ctx=ctx)
b, _ = api.prepare_view(raw[vs[0]], 0.002, 50, 1.0, 30, ctx=ctx)
ms, r = t(lambda: api.icp_align(a, b, 0.02, 50, mode=1, ctx=ctx))
print("icp resident  ms", ms, r["iterations"], r["ms"])
