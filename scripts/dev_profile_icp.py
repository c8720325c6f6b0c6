"""One resident point-to-plane alignment of the bench pair (for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from lowcost3dreconstruction_b200 import api
from lowcost3dreconstruction_b200._capi import HostCloud
mode = int(sys.argv[1]) if len(sys.argv) > 1 else 1
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
src, tgt = bench.load_pair(0)
ctx = api.Context(0)
n_t, c_t = api.normals(tgt, 30, ctx=ctx)
S, T = HostCloud(src), HostCloud(tgt, normal=n_t, curvature=c_t)
dS, dT = ctx.upload(S), ctx.upload(T)
for _ in range(reps):
    r = api.icp_align(dS, dT, 0.02, 50, mode=mode, ctx=ctx)
print(r["iterations"], r["ms"], ctx.grid_info())
