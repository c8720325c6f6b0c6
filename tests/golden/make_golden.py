"""Generates tests/golden/golden_small.npz from the CPU oracle (run from the repo root:
`python tests/golden/make_golden.py`).

The reference ships no golden vectors (parity unpinned) and cannot be run here
(PCL/Boost/Eigen/FLANN absent), so these fixtures pin OUR restatement: they are regression
anchors for the oracle (tests/test_oracle.py) and a second, committed target for the CUDA path
(tests/test_gpu_golden.py).  Inputs are seeded; regenerate only on an intentional semantic change.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from lowcost3dreconstruction_b200 import synth  # noqa: E402
from lowcost3dreconstruction_b200._capi import HostCloud  # noqa: E402
from oracle import oracle as orc  # noqa: E402


def main():
    tgt = synth.kinect_view(0, scale=0.1, backdrop="panel")
    src = synth.kinect_view(1, scale=0.1, backdrop="panel")
    out = dict(src=src, tgt=tgt)
    kt = orc.KdTree(tgt)
    out["nn_idx"], out["nn_d2"] = kt.nn(src, 0.02)
    out["knn_idx"], out["knn_d2"] = kt.knn(tgt, 12)
    nrm, curv = orc.normals(tgt, 12)
    out["normals"], out["curvature"] = nrm, curv
    T = HostCloud(tgt, normal=nrm, curvature=curv)
    for name, mode in (("p2p", 0), ("p2plane", 1)):
        r = orc.icp_align(src, T, 0.02, 50, mode=mode, dump_iteration=1)
        out[f"icp_{name}_T"] = r["transformation"]
        out[f"icp_{name}_meta"] = np.array([r["iterations"], r["state"], int(r["converged"])], dtype=np.int64)
        out[f"icp_{name}_fitness"] = np.array([r["fitness"]])
        out[f"icp_{name}_corr1"] = r["corr_index"]
    kept, md, st = orc.sor(tgt, 10, 1.0)
    out["sor_kept"], out["sor_mean"], out["sor_stats"] = kept, md, st
    v = orc.voxel_grid(HostCloud(tgt, normal=nrm, curvature=curv), 0.02)
    out["vox_xyz"], out["vox_of_point"], out["vox_normal"] = v["xyz"], v["voxel_of_point"], v["normal"]
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_small.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: getattr(v, "shape", None) for k, v in out.items()})


if __name__ == "__main__":
    main()
