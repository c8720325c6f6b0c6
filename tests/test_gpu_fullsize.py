"""GPU vs the CPU oracle at the BASELINE.json sizes (not scaled-down stand-ins): configs[1]
(307k x 307k point-to-plane, normals from the GPU k=30 pass), configs[0] (the reference's own
mode: ~200k point-to-point, 50-iteration cap), configs[3] (Kinect-v2 grid, ~1.5 M points per view,
with a top view registered against view 0), the in-process view pipeline, and the `-c` flag of
normal_estimation.  Bars as in tests/test_gpu_parity.py (north star): correspondence index sets
bit-exact, iteration counts / convergence state equal, transforms within 1e-4 rad and 1e-5 x
extent, fitness within 1e-5 relative."""
import os
import subprocess

import numpy as np
import pytest

from lowcost3dreconstruction_b200 import api, synth
from lowcost3dreconstruction_b200._capi import HostCloud
from oracle import oracle as orc
from tests import plyutil
from tests.test_gpu_parity import FIT_TOL, assert_transform_close

pytestmark = pytest.mark.gpu
BIN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "lowcost3dreconstruction_b200",
                   "tools", "bin")


def _compare_alignment(ctx, src, T, mode, dump_its, extent):
    o = None
    for it in dump_its:
        g = api.icp_align(src, T, 0.02, 50, mode=mode, dump_iteration=it, ctx=ctx)
        o = orc.icp_align(src, T, 0.02, 50, mode=mode, dump_iteration=it)
        assert np.array_equal(g["corr_index"], o["corr_index"]), f"correspondences differ at iteration {it}"
        assert np.array_equal(g["corr_dist2"][g["corr_index"] >= 0], o["corr_dist2"][o["corr_index"] >= 0])
    assert (g["iterations"], g["state"], g["converged"]) == (o["iterations"], o["state"], o["converged"])
    assert g["last_correspondences"] == o["last_correspondences"]
    assert_transform_close(g["transformation"], o["transformation"], extent)
    assert abs(g["fitness"] - o["fitness"]) <= FIT_TOL * o["fitness"]
    return g, o


def test_cfg2_full_size_point_to_plane_vs_oracle(ctx):
    tgt = synth.kinect_view(0, backdrop="full")
    src = synth.kinect_view(1, backdrop="full")
    assert len(tgt) > 300000 and len(src) > 300000
    nrm, curv = api.normals(tgt, 30, ctx=ctx)  # the pass that feeds ICP in the pipeline
    T = HostCloud(tgt, normal=nrm, curvature=curv)
    extent = float(np.ptp(tgt, axis=0).max())
    g, o = _compare_alignment(ctx, src, T, api.POINT_TO_PLANE, (0, 4), extent)
    assert g["iterations"] >= 5


def test_cfg1_full_size_point_to_point_vs_oracle(ctx):
    """pcl_tools/fine_registration.cpp:105 as the reference runs it: point-to-point, 50-iteration cap."""
    tgt = synth.kinect_view(0, backdrop="panel")
    src = synth.kinect_view(1, backdrop="panel")
    assert 150000 < len(tgt) < 250000
    extent = float(np.ptp(tgt, axis=0).max())
    g, o = _compare_alignment(ctx, src, HostCloud(tgt), api.POINT_TO_POINT, (0, 10), extent)
    # PCL's own Umeyama sums run in float32.  Summing ~200k float32 terms carries a
    # relative error of ~1e-4 that depends on the summation order (Eigen vectorises it; ours is
    # sequential), so against the oracle's float32-sum mode the bars are the noise floor of that
    # arithmetic, 10x the north-star bars — the fp64-sum comparison above is the tight one.  The
    # iteration count and the convergence state still have to agree.
    o32 = orc.icp_align(src, tgt, 0.02, 50, mode=0, umeyama_f32=True)
    assert g["iterations"] == o32["iterations"] and g["state"] == o32["state"]
    assert np.abs(g["transformation"] - o32["transformation"]).max() <= 1e-4 * extent
    assert abs(g["fitness"] - o32["fitness"]) <= 1e-3 * o32["fitness"]


def test_cfg4_kinect_v2_1p5m_with_top_view(ctx):
    """BASELINE configs[3]: Kinect-v2 frustum on a 1344 x 1113 grid (~1.5 M points per view), 15 deg
    steps; pair 1 -> 0 and a top view (pitched 40 deg, pre-placed like centroid_align does at
    scripts/alignment.sh:118-119, with a residual error) registered against view 0."""
    v0 = synth.kinect_v2_sr_view(0)
    assert len(v0) > 1_400_000
    nrm, curv = api.normals(v0, 30, ctx=ctx)
    T = HostCloud(v0, normal=nrm, curvature=curv)
    extent = float(np.ptp(v0, axis=0).max())
    # turntable pair, pre-aligned with the rotate_align prior up to a residual
    v1 = synth.kinect_v2_sr_view(1)
    resid = synth.rigid(0.5, -0.7, 0.3, [0.003, -0.002, 0.004])
    src = synth.apply_transform(resid @ synth.turntable_prior(1, 15.0), v1)
    g, o = _compare_alignment(ctx, src, T, api.POINT_TO_PLANE, (0,), extent)
    Rerr = g["transformation"][:3, :3].astype(np.float64) @ resid[:3, :3]
    assert np.degrees(np.arccos(np.clip((np.trace(Rerr) - 1) / 2, -1, 1))) < 0.2  # undoes the residual
    # top view against view 0
    top = synth.kinect_v2_sr_view(0, tilt_deg=40.0)
    src = synth.apply_transform(resid @ synth.tilt_motion(40.0), top)
    g = api.icp_align(src, T, 0.02, 50, mode=api.POINT_TO_PLANE, ctx=ctx)
    o = orc.icp_align(src, T, 0.02, 50, mode=1)
    assert (g["iterations"], g["state"]) == (o["iterations"], o["state"])
    assert_transform_close(g["transformation"], o["transformation"], extent)
    assert abs(g["fitness"] - o["fitness"]) <= FIT_TOL * o["fitness"]


def test_prepare_view_equals_stage_by_stage(ctx):
    """lc3d_prepare_view (VoxelGrid -> SOR -> normals on the device) is bit-identical to calling the
    three stage entry points through host buffers, and feeds the resident ICP."""
    raw = synth.kinect_view(2, step_deg=10.0, backdrop="none")
    d, cnt = api.prepare_view(raw, 0.002, 50, 1.0, 30, ctx=ctx)
    vox = api.voxel_grid(raw, 0.002, ctx=ctx)["xyz"]
    kept, _, _ = api.sor(vox, 50, 1.0, ctx=ctx)
    pts = vox[kept]
    nrm, curv = api.normals(pts, 30, ctx=ctx)
    xyz, dn, dc = d.download()
    assert cnt == (len(vox), len(pts), len(pts))
    assert np.array_equal(xyz, pts) and np.array_equal(dn, nrm, equal_nan=True) and np.array_equal(dc, curv, equal_nan=True)
    # stages can be skipped
    d2, cnt2 = api.prepare_view(raw, 0.0, 0, 1.0, 0, ctx=ctx)
    x2, n2, _ = d2.download()
    assert cnt2 == (len(raw),) * 3 and np.array_equal(x2, raw) and n2 is None
    # resident ICP on prepared views == host-buffer ICP on the same data
    raw1 = synth.kinect_view(3, step_deg=10.0, backdrop="none")
    prior = synth.turntable_prior(1, 10.0)
    d1, _ = api.prepare_view(synth.apply_transform(prior, raw1), 0.002, 50, 1.0, 30, ctx=ctx)
    r = api.icp_align(d1, d, 0.02, 50, mode=api.POINT_TO_PLANE, ctx=ctx)
    x1, n1, c1 = d1.download()
    rh = api.icp_align(HostCloud(x1, normal=n1, curvature=c1), HostCloud(xyz, normal=dn, curvature=dc), 0.02, 50,
                       mode=api.POINT_TO_PLANE, ctx=ctx)
    assert r["iterations"] == rh["iterations"] and np.array_equal(r["transformation"], rh["transformation"])
    assert r["fitness"] == rh["fitness"]


def test_normal_estimation_centroid_flag(tmp_path, ctx):
    """normal_estimation -c (pcl_tools/normal_estimation.cpp:98-102,112-118): viewpoint = float32
    centroid of the cloud, then ALL normals negated (centroid XOR reverse)."""
    pts = synth.kinect_view(0, scale=0.25, backdrop="none")
    src, out = str(tmp_path / "in.ply"), str(tmp_path / "out.ply")
    plyutil.write_capture_ascii(src, pts)
    p = subprocess.run([os.path.join(BIN, "normal_estimation"), "-i", src, "-o", out, "-n", "15", "-c"],
                       capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr
    loaded = plyutil.read_pcl_binary(out)
    xyz = np.array(loaded["xyz"])
    c = api.centroid(xyz, ctx=ctx)
    assert np.array_equal(c[:3], orc.centroid(xyz)[:3])
    nrm, curv = api.normals(xyz, 15, viewpoint=c[:3], ctx=ctx)
    assert np.array_equal(loaded["normal"], -nrm) and np.array_equal(loaded["curvature"], curv)
    # -c -r: the two negations cancel
    p = subprocess.run([os.path.join(BIN, "normal_estimation"), "-i", src, "-o", out, "-n", "15", "-c", "-r"],
                       capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr
    assert np.array_equal(plyutil.read_pcl_binary(out)["normal"], nrm)
    # pointing away from the centroid: outward for a convex-ish object
    inner = np.einsum("ij,ij->i", -nrm.astype(np.float64), xyz.astype(np.float64) - c[:3].astype(np.float64))
    assert np.nanmean(inner > 0) > 0.95


def test_native_chain_executor_matches_stage_calls(ctx):
    """lc3d_chain_run (the pair block as a task graph on the library's own host threads) returns, for
    every pair, the bits of lc3d_prepare_view + lc3d_icp_align_resident called one after the other."""
    from lowcost3dreconstruction_b200 import chain
    views = [synth.apply_transform(synth.turntable_prior(v, 10.0), synth.kinect_view(v, step_deg=10.0, backdrop="none", scale=0.5))
             for v in range(5)]
    kw = dict(leaf_size=0.004, sor_mean_k=30, sor_stddev_mul=1.0, normals_k=20)
    prepared = [api.prepare_view(v, kw["leaf_size"], kw["sor_mean_k"], kw["sor_stddev_mul"], kw["normals_k"], ctx=ctx)
                for v in views]
    ref = [api.icp_align(prepared[i + 1][0], prepared[i][0], 0.02, 50, mode=api.POINT_TO_PLANE, ctx=ctx) for i in range(4)]
    nat = chain.NativeChain(0, prepare_threads=3, align_threads=2)
    try:
        for warm in (True, False, False):
            res, npts = nat.run(views, max_correspondence_distance=0.02, max_iterations=50, mode=api.POINT_TO_PLANE,
                                warm=warm, **kw)
            assert npts == [p[1][2] for p in prepared]
            for r, g in zip(ref, res):
                assert np.array_equal(r["transformation"], g["transformation"]) and r["fitness"] == g["fitness"]
                assert r["iterations"] == g["iterations"] and r["state"] == g["state"]
        with pytest.raises(RuntimeError):
            nat.run(views, mode=api.POINT_TO_PLANE)  # point-to-plane without normals: error surfaces, nothing hangs
    finally:
        nat.close()
    for p in prepared:
        p[0].free()
