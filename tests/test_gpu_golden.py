"""CUDA path vs the committed golden fixtures (tests/golden/golden_small.npz, made by
tests/golden/make_golden.py from the oracle) and a full-size property test."""
import os

import numpy as np
import pytest

from lowcost3dreconstruction_b200 import api, synth
from lowcost3dreconstruction_b200._capi import HostCloud

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_small.npz")


def test_cuda_reproduces_golden_fixtures(ctx):
    g = np.load(GOLDEN)
    src, tgt = g["src"], g["tgt"]
    idx, d2 = api.nn(tgt, src, 0.02, ctx=ctx)
    assert np.array_equal(idx, g["nn_idx"]) and np.array_equal(d2, g["nn_d2"])
    ki, kd = api.knn(tgt, 12, ctx=ctx)
    assert np.array_equal(ki, g["knn_idx"]) and np.array_equal(kd, g["knn_d2"])
    T = HostCloud(tgt, normal=g["normals"], curvature=g["curvature"])
    for name, mode in (("p2p", 0), ("p2plane", 1)):
        r = api.icp_align(src, T, 0.02, 50, mode=mode, dump_iteration=1, ctx=ctx)
        assert [r["iterations"], r["state"], int(r["converged"])] == g[f"icp_{name}_meta"].tolist()
        assert np.abs(r["transformation"] - g[f"icp_{name}_T"]).max() < 1e-6
        assert np.array_equal(r["corr_index"], g[f"icp_{name}_corr1"])
        assert abs(r["fitness"] - g[f"icp_{name}_fitness"][0]) <= 1e-5 * g[f"icp_{name}_fitness"][0]
    kept, md, st = api.sor(tgt, 10, 1.0, ctx=ctx)
    assert np.array_equal(kept, g["sor_kept"]) and np.array_equal(md, g["sor_mean"])
    v = api.voxel_grid(T, 0.02, ctx=ctx)
    assert np.array_equal(v["voxel_of_point"], g["vox_of_point"]) and np.array_equal(v["xyz"], g["vox_xyz"])
    nrm, curv = api.normals(tgt, 12, ctx=ctx)
    ang = np.linalg.norm(np.cross(nrm.astype(np.float64), g["normals"].astype(np.float64)), axis=1)
    assert np.nanmax(ang) < 2e-3 and np.nanmedian(ang) < 1e-6


def test_full_size_properties(ctx):
    """BASELINE full size (307k x 307k): size-independent properties instead of the oracle."""
    tgt = synth.kinect_view(0, backdrop="full")
    src = synth.kinect_view(1, backdrop="full")
    assert len(tgt) > 300000
    # NN of a cloud against itself is the identity with zero distance
    idx, d2 = api.nn(tgt, tgt, 0.0, ctx=ctx)
    assert np.array_equal(idx, np.arange(len(tgt))) and not d2.any()
    # kNN: sorted, self first, k-th distance consistent with a radius count on a sample
    ki, kd = api.knn(tgt, 8, ctx=ctx)
    assert np.all(np.diff(kd, axis=1) >= 0) and np.array_equal(ki[:, 0], np.arange(len(tgt)))
    # ICP: result is a rigid transform, registered cloud == T * source, fitness == mean NN d2
    nrm, curv = api.normals(tgt, 30, ctx=ctx)
    r = api.icp_align(src, HostCloud(tgt, normal=nrm), 0.02, 50, mode=1, want_registered=True, ctx=ctx)
    R = r["transformation"][:3, :3].astype(np.float64)
    assert np.allclose(R @ R.T, np.eye(3), atol=1e-5) and abs(np.linalg.det(R) - 1) < 1e-5
    reg, _ = api.transform(src, r["transformation"], ctx=ctx)
    assert np.array_equal(reg, r["registered_xyz"])
    _, d2 = api.nn(tgt, reg, 0.0, ctx=ctx)
    assert abs(d2.astype(np.float64).mean() - r["fitness"]) <= 1e-9 * max(r["fitness"], 1e-30) + 1e-12
    # recovered motion is the 5 degree turntable step
    ang = np.degrees(np.arctan2(R[2, 0], R[0, 0]))
    assert abs(ang - 5.0) < 0.3
    # voxel grid: every input point lands in exactly one voxel; SOR(negative) complements SOR
    v = api.voxel_grid(tgt, 0.002, ctx=ctx)
    assert v["voxel_of_point"].min() == 0 and v["voxel_of_point"].max() == len(v["xyz"]) - 1
    k1, _, _ = api.sor(tgt, 50, 1.0, ctx=ctx)
    k2, _, _ = api.sor(tgt, 50, 1.0, negative=True, ctx=ctx)
    assert len(k1) + len(k2) == len(tgt) and not np.intersect1d(k1, k2).size


def test_cuda_matches_independent_golden(ctx):
    """The fixture of tests/golden/make_golden_independent.py (numpy / scipy, no oracle code):
    correspondences bit-exact at two iterations, iteration counts and convergence states equal —
    including the TRANSFORM and ABS_MSE exits — transforms / fitness within the north-star bars."""
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_independent.npz"))
    src, tgt = g["src"], g["tgt"]
    idx, d2 = api.nn(tgt, src, 0.02, ctx=ctx)
    assert np.array_equal(idx, g["nn_idx"]) and np.array_equal(d2[idx >= 0], g["nn_d2"][idx >= 0])
    T = HostCloud(tgt, normal=g["normals"], curvature=g["curvature"])
    extent = float(np.ptp(tgt, axis=0).max())
    for name, mode in (("p2p", 0), ("p2plane", 1)):
        for it in (0, 3):
            r = api.icp_align(src, T, 0.02, 50, mode=mode, dump_iteration=it, ctx=ctx)
            assert np.array_equal(r["corr_index"], g[f"icp_{name}_corr{it}"]), (name, it)
        assert [r["iterations"], r["state"]] == g[f"icp_{name}_meta"].tolist()
        assert np.abs(r["transformation"] - g[f"icp_{name}_T"])[:3, :3].max() < 1e-5
        assert np.abs(r["transformation"] - g[f"icp_{name}_T"])[:3, 3].max() < 1e-5 * extent
        assert abs(r["fitness"] - g[f"icp_{name}_fitness"][0]) <= 1e-5 * r["fitness"]
    r = api.icp_align(src, T, 0.02, 50, transformation_epsilon=1e-5, euclidean_fitness_epsilon=0.0, mode=1, ctx=ctx)
    assert [r["iterations"], r["state"]] == g["icp_transform_exit_meta"].tolist() and r["state"] == 2
    assert np.abs(r["transformation"] - g["icp_transform_exit_T"]).max() < 1e-5
    r = api.icp_align(g["near"], tgt, 0.02, 50, transformation_epsilon=0.0, euclidean_fitness_epsilon=0.0, mode=0,
                      ctx=ctx)
    assert [r["iterations"], r["state"]] == g["icp_abs_mse_exit_meta"].tolist() and r["state"] == 3
    kept, mean, _ = api.sor(tgt, 10, 1.0, ctx=ctx)
    assert np.array_equal(kept, g["sor_kept"]) and np.array_equal(mean, g["sor_mean"])
    v = api.voxel_grid(tgt, 0.02, ctx=ctx)
    assert np.array_equal(v["voxel_of_point"], g["vox_of_point"]) and np.abs(v["xyz"] - g["vox_xyz"]).max() < 1e-6
