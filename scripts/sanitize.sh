#!/bin/bash
# compute-sanitizer passes over a small host-buffer + resident alignment and the filters (developer tool)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cat > /tmp/san_job.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np
from lowcost3dreconstruction_b200 import api, synth
from lowcost3dreconstruction_b200._capi import HostCloud
ctx = api.Context(0)
tgt = synth.kinect_view(0, scale=0.25, backdrop="panel")
src = synth.kinect_view(1, scale=0.25, backdrop="panel")
nrm, curv = api.normals(tgt, 20, ctx=ctx)
T = HostCloud(tgt, normal=nrm, curvature=curv)
for mode in (1, 0):
    g = api.icp_align(src, T, 0.02, 50, mode=mode, ctx=ctx)
    print("host", mode, g["iterations"], g["last_correspondences"], g["fitness"])
dS, dT = ctx.upload(HostCloud(src)), ctx.upload(T)
for mode in (1, 0):
    g = api.icp_align(dS, dT, 0.02, 50, mode=mode, ctx=ctx)
    print("resident", mode, g["iterations"], g["last_correspondences"], g["fitness"])
kept, _, _ = api.sor(tgt, 20, 1.0, ctx=ctx)
v = api.voxel_grid(tgt, 0.01, ctx=ctx)
d, cnt = api.prepare_view(tgt, 0.004, 20, 1.0, 20, ctx=ctx)
print("filters", len(kept), len(v["xyz"]), cnt)
PY
for tool in ${TOOLS:-initcheck racecheck memcheck}; do
  echo "== $tool"
  timeout 900 compute-sanitizer --tool $tool --print-limit 30 python /tmp/san_job.py > gpurun_out/san_$tool.log 2>&1
  grep -E "ERROR SUMMARY|host |resident |filters" gpurun_out/san_$tool.log | tail -8
  grep -E "Uninitialized|Race|Invalid|hazard" gpurun_out/san_$tool.log | sort | uniq -c | sort -rn | head -12
  grep -A12 -m1 -E "Uninitialized|hazard|Invalid" gpurun_out/san_$tool.log | head -30
done
