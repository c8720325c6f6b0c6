"""Same-box A/B of liblc3d.so variants on the bench pair (developer tool, not a test).

  python scripts/ab_icp.py [--stats] [--p2p] lib_a.so lib_b.so ...

Every variant runs in its own process (LC3D_LIB is read at import): resident point-to-plane
alignment of the 307k bench pair, best / median of the phase timings the library reports, and a
checksum of the result (iterations, state, fitness, transform) so that a variant that changes
the answer is visible immediately."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r"""
import sys, os, hashlib
sys.path.insert(0, %(root)r)
import numpy as np
import bench
from lowcost3dreconstruction_b200 import api
from lowcost3dreconstruction_b200._capi import HostCloud
src, tgt = bench.load_pair(0)
ctx = api.Context(0)
n_t, c_t = api.normals(tgt, 30, ctx=ctx)
dS, dT = ctx.upload(HostCloud(src)), ctx.upload(HostCloud(tgt, normal=n_t, curvature=c_t))
print("  inputs", hashlib.sha1(src.tobytes()).hexdigest()[:10], hashlib.sha1(tgt.tobytes()).hexdigest()[:10],
      "normals", hashlib.sha1(n_t.tobytes()).hexdigest()[:10], flush=True)
modes = %(modes)r
for mode in modes:
    rows, seen = [], set()
    for rep in range(%(reps)d):
        r = api.icp_align(dS, dT, 0.02, 50, mode=mode, ctx=ctx)
        rows.append([r['ms']['index'], r['ms']['loop'], r['ms']['fitness'], r['ms']['total']])
        seen.add((r['iterations'], r['last_correspondences'], r['fitness'], r['transformation'].tobytes()))
    if len(seen) != 1:
        print(f"  !! mode {mode}: {len(seen)} DISTINCT outcomes in {%(reps)d} repetitions", flush=True)
    a = np.array(rows[3:])
    h = hashlib.sha1(r['transformation'].tobytes()).hexdigest()[:10]
    print(f"  mode {mode}: it {r['iterations']} st {r['state']} corr {r['last_correspondences']} fit {r['fitness']:.9e} T#{h} | "
          f"min index {a[:,0].min():.4f} loop {a[:,1].min():.4f} fitness {a[:,2].min():.4f} total {a[:,3].min():.4f} | "
          f"median loop {np.median(a[:,1]):.4f} total {np.median(a[:,3]):.4f}", flush=True)
"""


def main():
    args = sys.argv[1:]
    stats = "--stats" in args
    modes = [0, 1] if "--both" in args else ([0] if "--p2p" in args else [1])
    libs = [a for a in args if not a.startswith("--")]
    for lib in libs:
        env = dict(os.environ)
        if lib != "default":
            env["LC3D_LIB"] = os.path.abspath(lib)
        print(f"== {lib}", flush=True)
        code = CHILD % {"root": ROOT, "modes": modes, "reps": 23}
        subprocess.run([sys.executable, "-c", code], env=env, check=False)
        if stats:
            env["LC3D_STATS"] = "1"
            code = CHILD % {"root": ROOT, "modes": modes, "reps": 4}
            p = subprocess.run([sys.executable, "-c", code], env=env, check=False, capture_output=True, text=True)
            lines = [l for l in p.stderr.splitlines() if "[lc3d stats]" in l]
            # the last alignment's block only
            last = max((i for i, l in enumerate(lines) if "fitness searched" in l), default=0)
            print("\n".join(l[:260] for l in lines[last:]), flush=True)


if __name__ == "__main__":
    main()
