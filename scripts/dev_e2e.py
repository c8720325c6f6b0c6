"""Developer check: where the host-buffer (e2e) call spends its time."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from lowcost3dreconstruction_b200 import api
from lowcost3dreconstruction_b200._capi import HostCloud
src, tgt = bench.load_pair(0)
ctx = api.Context(0)
n_t, c_t = api.normals(tgt, 30, ctx=ctx)
n_s, c_s = api.normals(src, 30, ctx=ctx)
def pinned(a):
    t_ = torch.from_numpy(np.ascontiguousarray(a)).pin_memory(); return t_, t_.numpy()
keep = [pinned(x) for x in (src, n_s, tgt, n_t)]
Sp = HostCloud(keep[0][1], normal=keep[1][1]); Tp = HostCloud(keep[2][1], normal=keep[3][1])
Su = HostCloud(src, normal=n_s); Tu = HostCloud(tgt, normal=n_t)
for name, S, T in (("pinned", Sp, Tp), ("pageable", Su, Tu)):
    for want in (True, False):
        for _ in range(3):
            api.icp_align(S, T, 0.02, 50, mode=1, want_registered=want, ctx=ctx)
        t0 = time.perf_counter(); N = 10
        for _ in range(N):
            r = api.icp_align(S, T, 0.02, 50, mode=1, want_registered=want, ctx=ctx)
        wall = (time.perf_counter() - t0) / N * 1e3
        print(f"{name:9s} registered={want}: wall {wall:.3f} ms; device total {r['ms']['total']:.3f} = upload {r['ms']['upload']:.3f} + index {r['ms']['index']:.3f} + loop {r['ms']['loop']:.3f} + fitness {r['ms']['fitness']:.3f} + download {r['ms']['download']:.3f}")
# raw PCIe rates
a = torch.empty(64 << 20, dtype=torch.uint8).pin_memory(); d = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
for _ in range(3): d.copy_(a, non_blocking=True); torch.cuda.synchronize()
t0 = time.perf_counter(); d.copy_(a, non_blocking=True); torch.cuda.synchronize(); h2d = 64 / 1024 / (time.perf_counter() - t0)
t0 = time.perf_counter(); a.copy_(d, non_blocking=True); torch.cuda.synchronize(); d2h = 64 / 1024 / (time.perf_counter() - t0)
print(f"PCIe pinned: H2D {h2d:.1f} GB/s, D2H {d2h:.1f} GB/s")
