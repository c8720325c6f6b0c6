"""Developer perf probe (not a test): phase timings of the hot path at full Kinect size."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lowcost3dreconstruction_b200 import api, synth
from lowcost3dreconstruction_b200._capi import HostCloud

backdrop = sys.argv[1] if len(sys.argv) > 1 else "full"
step = float(sys.argv[2]) if len(sys.argv) > 2 else 5.0
t0 = time.time()
c0 = synth.kinect_view(0, step_deg=step, backdrop=backdrop)
c1 = synth.kinect_view(1, step_deg=step, backdrop=backdrop)
print("clouds", c0.shape, c1.shape, "render s", round(time.time() - t0, 1), flush=True)
ctx = api.Context(0)
for rep in range(2):
    t = time.time(); n0, k0 = api.normals(c0, 30, ctx=ctx); dt = time.time() - t
    print(f"normals k=30 wall {dt*1e3:.2f} ms", flush=True)
t = time.time(); n1, k1 = api.normals(c1, 30, ctx=ctx)
for rep in range(2):
    t = time.time(); kept, md, st = api.sor(c0, 50, 1.0, ctx=ctx); dt = time.time() - t
    print(f"sor k=50 wall {dt*1e3:.2f} ms kept {len(kept)}", flush=True)
for rep in range(2):
    t = time.time(); v = api.voxel_grid(c0, 0.002, ctx=ctx); dt = time.time() - t
    print(f"voxel 2mm wall {dt*1e3:.2f} ms -> {len(v['xyz'])}", flush=True)
S = HostCloud(c1, normal=n1, curvature=k1); T = HostCloud(c0, normal=n0, curvature=k0)
ds, dt_ = ctx.upload(S), ctx.upload(T)
for mode in (0, 1):
    for rep in range(3):
        l0 = ctx.launch_count
        r = api.icp_align(ds, dt_, 0.02, 50, mode=mode, ctx=ctx)
        print(f"mode={mode} it={r['iterations']} st={r['state']} corr={r['last_correspondences']} fit={r['fitness']:.4e} "
              f"index {r['ms']['index']:.3f} loop {r['ms']['loop']:.3f} ms ({r['ms']['loop']/max(r['iterations'],1)*1e3:.1f} us/it) "
              f"fitness {r['ms']['fitness']:.3f} total {r['ms']['total']:.3f} launches {ctx.launch_count-l0}", flush=True)
    t = time.time(); r = api.icp_align(S, T, 0.02, 50, mode=mode, ctx=ctx); dt = time.time() - t
    print(f"   host-buffer call wall {dt*1e3:.2f} ms; upload {r['ms']['upload']:.3f} total {r['ms']['total']:.3f}")
print(r["transformation"]); print(synth.turntable_motion(step))
