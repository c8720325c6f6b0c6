#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for b in 26 27 28 29; do LC3D_MAX_CELLS_LOG2=$b timeout 600 python scripts/dev_large.py 5e6 2>&1 | grep -v "^normals\|^sor\|^voxel" | cut -c1-330; done
timeout 600 python scripts/dev_large.py 5e6 2>&1 | cut -c1-330 | tee gpurun_out/r02_cfg5_5M.log
echo "== no-op launches: max_iter 9 vs 50"
python - <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import bench
from lowcost3dreconstruction_b200 import api
from lowcost3dreconstruction_b200._capi import HostCloud
src, tgt = bench.load_pair(0)
ctx = api.Context(0)
n_t, c_t = api.normals(tgt, 30, ctx=ctx)
dS, dT = ctx.upload(HostCloud(src)), ctx.upload(HostCloud(tgt, normal=n_t, curvature=c_t))
for mi in (50, 9, 50, 9):
    for _ in range(6):
        r = api.icp_align(dS, dT, 0.02, mi, mode=1, ctx=ctx)
    print(mi, r["iterations"], r["state"], r["ms"])
PY
