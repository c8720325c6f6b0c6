"""MVS-scale clouds (BASELINE configs[4]: 5M x 5M shuffled, point-to-plane): feeding passes, ICP in
both modes, error against the known motion.  Output kept under profiles/ (r02_cfg5_*.log)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lowcost3dreconstruction_b200 import api, synth
from lowcost3dreconstruction_b200._capi import HostCloud
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 5_000_000
ctx = api.Context(0)
tgt = synth.surface_samples(n, seed=5)
T = synth.rigid(0.0, 3.0, 0.0, [0.005, 0.0, 0.0])
src = synth.apply_transform(np.linalg.inv(T), synth.surface_samples(n, seed=6))
print(f"n = {n}, LC3D_MAX_CELLS_LOG2 = {os.environ.get('LC3D_MAX_CELLS_LOG2', 'default')}", flush=True)
for rep in range(2):
    t = time.time(); nrm, curv = api.normals(tgt, 30, ctx=ctx); dt = time.time() - t
print(f"normals k=30 (host call): {dt:.3f} s", ctx.grid_info(), flush=True)
for rep in range(2):
    t = time.time(); kept, md, st = api.sor(tgt, 50, 1.0, ctx=ctx); dt = time.time() - t
print(f"sor k=50 (host call): {dt:.3f} s kept {len(kept)}", flush=True)
for rep in range(2):
    t = time.time(); v = api.voxel_grid(tgt, 0.001, ctx=ctx); dt = time.time() - t
print(f"voxel 1 mm (host call): {dt:.3f} s -> {len(v['xyz'])}", flush=True)
S, Tg = HostCloud(src), HostCloud(tgt, normal=nrm)
dS, dT = ctx.upload(S), ctx.upload(Tg)
for mode in (1, 0):
    for rep in range(3):
        r = api.icp_align(dS, dT, 0.02, 50, mode=mode, ctx=ctx)
    it = max(r["iterations"], 1)
    print(f"mode {mode}: it={r['iterations']} state={r['state']} corr={r['last_correspondences']} fit={r['fitness']:.3e} "
          f"ms={ {k: round(v, 3) for k, v in r['ms'].items()} } -> {it / (r['ms']['total'] * 1e-3):.0f} it/s whole alignment, "
          f"{r['ms']['loop'] / it * 1e3:.0f} us/iteration = {64.0 * n / (r['ms']['loop'] / it * 1e-3) / 1e9:.0f} GB/s algorithmic",
          ctx.grid_info(), flush=True)
    err = np.abs(r["transformation"].astype(np.float64) - T).max()
    print("   max |T - T_true| =", err, flush=True)
