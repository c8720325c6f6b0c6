"""Repeats the resident bench alignment and counts distinct outcomes (developer tool)."""
import sys, os, hashlib, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from lowcost3dreconstruction_b200 import api
from lowcost3dreconstruction_b200._capi import HostCloud
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
src, tgt = bench.load_pair(0)
ctx = api.Context(0)
n_t, c_t = api.normals(tgt, 30, ctx=ctx)
print("normals hash", hashlib.sha1(n_t.tobytes()).hexdigest()[:10], "src", hashlib.sha1(src.tobytes()).hexdigest()[:10],
      "tgt", hashlib.sha1(tgt.tobytes()).hexdigest()[:10])
dS, dT = ctx.upload(HostCloud(src)), ctx.upload(HostCloud(tgt, normal=n_t, curvature=c_t))
for mode in (1, 0):
    seen = collections.Counter()
    for rep in range(reps if mode else reps // 4):
        r = api.icp_align(dS, dT, 0.02, 50, mode=mode, ctx=ctx)
        key = (r["iterations"], r["last_correspondences"], f"{r['fitness']:.12e}", hashlib.sha1(r["transformation"].tobytes()).hexdigest()[:10])
        seen[key] += 1
    print("mode", mode, "distinct outcomes:", len(seen))
    for k, v in seen.most_common(6):
        print("   ", v, k)
