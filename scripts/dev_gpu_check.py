"""Developer GPU check (not a test): quick parity + timing printout."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lowcost3dreconstruction_b200 import api, synth
from oracle import oracle as orc

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.5
c0 = synth.kinect_view(0, scale=scale, backdrop="panel")
c1 = synth.kinect_view(1, scale=scale, backdrop="panel")
print("clouds", c0.shape, c1.shape, flush=True)
ctx = api.Context(0)
kt = orc.KdTree(c0)
for md in (0.0, 0.02, 0.005):
    oi, od = kt.nn(c1, md)
    gi, gd = api.nn(c0, c1, md, ctx=ctx)
    mism = np.nonzero(gi != oi)[0]
    print(f"nn max_dist={md}: idx mismatches {len(mism)} / {len(oi)}; d2 equal where idx equal:",
          np.array_equal(gd[gi == oi], od[gi == oi]), " matched frac", (oi >= 0).mean(), flush=True)
    if len(mism):
        print("   sample", mism[:5], gi[mism[:5]], oi[mism[:5]], gd[mism[:5]], od[mism[:5]])
for mode in (0,):
    for it in (0, 3, 10):
        o = orc.icp_align(c1, c0, 0.02, 50, mode=mode, dump_iteration=it)
        g = api.icp_align(c1, c0, 0.02, 50, mode=mode, dump_iteration=it, ctx=ctx)
        print(f"icp mode={mode} dump_it={it}: oracle it={o['iterations']} st={o['state']} fit={o['fitness']:.6e} | gpu it={g['iterations']} st={g['state']} fit={g['fitness']:.6e}")
        print("   corr mismatches", (o['corr_index'] != g['corr_index']).sum(), "of", len(c1),
              " max|dT|", np.abs(o['transformation'] - g['transformation']).max(), g['ms'], flush=True)
print(g['transformation'])
