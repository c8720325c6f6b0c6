"""GPU parity tests: CUDA path (through the C ABI) vs the CPU oracle on seeded inputs.

Bars (BASELINE.json north star): correspondence / neighbour index sets bit-exact (exact
distance ties are resolved identically here: lower index), voxel / outlier masks bit-exact,
final transforms within 1e-4 rad and 1e-5 x extent, fitness within 1e-5 relative.
"""
import numpy as np
import pytest

from lowcost3dreconstruction_b200 import api, synth
from lowcost3dreconstruction_b200._capi import HostCloud
from oracle import oracle as orc

pytestmark = pytest.mark.gpu

ROT_TOL = 1e-4      # rad
TRANS_TOL = 1e-5    # x cloud extent
FIT_TOL = 1e-5      # relative


def rot_angle(Ta, Tb):
    """Relative rotation angle, from the antisymmetric part (arccos of the trace loses half
    the digits near zero and float32 matrices are only orthonormal to ~1e-7)."""
    R = Ta[:3, :3].astype(np.float64) @ Tb[:3, :3].astype(np.float64).T
    w = 0.5 * np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    s = float(np.linalg.norm(w))
    c = (np.trace(R) - 1.0) / 2.0
    return float(np.arctan2(s, c))


def assert_transform_close(Tg, To, extent):
    assert rot_angle(Tg, To) <= ROT_TOL
    assert np.abs(Tg[:3, 3].astype(np.float64) - To[:3, 3]).max() <= TRANS_TOL * extent


@pytest.fixture(scope="module")
def pair():
    tgt = synth.kinect_view(0, scale=0.4, backdrop="panel")
    src = synth.kinect_view(1, scale=0.4, backdrop="panel")
    return src, tgt


@pytest.fixture(scope="module")
def pair_normals(pair, ctx):
    src, tgt = pair
    n_t, c_t = orc.normals(tgt, 20)
    n_s, c_s = orc.normals(src, 20)
    return HostCloud(src, normal=n_s, curvature=c_s), HostCloud(tgt, normal=n_t, curvature=c_t)


# ------------------------------------------------------------------ 1-NN / correspondences

@pytest.mark.parametrize("max_dist", [0.0, 0.02, 0.004])
def test_nn_bit_exact(ctx, pair, max_dist):
    src, tgt = pair
    oi, od = orc.KdTree(tgt).nn(src, max_dist)
    gi, gd = api.nn(tgt, src, max_dist, ctx=ctx)
    assert np.array_equal(gi, oi)
    assert np.array_equal(gd, od)  # float32 squared distances, bit for bit


def test_nn_queries_far_outside_grid(ctx):
    rng = np.random.default_rng(3)
    tgt = rng.uniform(-1, 1, (5000, 3)).astype(np.float32)
    q = np.concatenate([rng.uniform(-30, 30, (2000, 3)), rng.uniform(-1, 1, (500, 3))]).astype(np.float32)
    oi, od = orc.KdTree(tgt).nn(q)
    gi, gd = api.nn(tgt, q, ctx=ctx)
    assert np.array_equal(gi, oi) and np.array_equal(gd, od)


def test_nn_degenerate_inputs(ctx):
    one = np.array([[0.5, -1.0, 2.0]], dtype=np.float32)
    q = np.array([[0, 0, 0], [1, 1, 1]], dtype=np.float32)
    gi, gd = api.nn(one, q, ctx=ctx)
    assert gi.tolist() == [0, 0]
    # all-identical points: tie -> lowest index
    same = np.tile(one, (100, 1))
    gi, gd = api.nn(same, q, ctx=ctx)
    assert gi.tolist() == [0, 0]
    # duplicates + NaN points in the target are skipped
    t = np.array([[0, 0, 0], [np.nan, 0, 0], [1, 1, 1], [1, 1, 1]], dtype=np.float32)
    gi, gd = api.nn(t, q, ctx=ctx)
    assert gi.tolist() == [0, 2]
    # planar / collinear clouds (zero extent in some dimension)
    rng = np.random.default_rng(0)
    plane = rng.uniform(0, 1, (3000, 3)).astype(np.float32)
    plane[:, 2] = 0.25
    qq = rng.uniform(-0.5, 1.5, (1000, 3)).astype(np.float32)
    oi, od = orc.KdTree(plane).nn(qq)
    gi, gd = api.nn(plane, qq, ctx=ctx)
    assert np.array_equal(gi, oi) and np.array_equal(gd, od)


def test_nn_volumetric_random(ctx):
    rng = np.random.default_rng(11)
    tgt = rng.normal(0, 1, (40000, 3)).astype(np.float32)
    q = rng.normal(0, 1.3, (30000, 3)).astype(np.float32)
    for md in (0.0, 0.05):
        oi, od = orc.KdTree(tgt).nn(q, md)
        gi, gd = api.nn(tgt, q, md, ctx=ctx)
        assert np.array_equal(gi, oi) and np.array_equal(gd, od)


def test_nn_anisotropic_grid_shapes(ctx):
    """The index slices its cells 8x along x (GridDev::xs) and backs off when x is the long axis:
    exercise the subdivision limits with clouds stretched along each axis."""
    rng = np.random.default_rng(21)
    for axis, length in ((0, 100.0), (1, 100.0), (2, 100.0), (0, 0.0)):
        tgt = rng.normal(0, 0.01, (12000, 3)).astype(np.float32)
        tgt[:, axis] = rng.uniform(0, length, len(tgt)).astype(np.float32) if length else np.float32(0.5)
        q = tgt[rng.choice(len(tgt), 6000)] + rng.normal(0, 0.02, (6000, 3)).astype(np.float32)
        for md in (0.0, 0.03):
            oi, od = orc.KdTree(tgt).nn(q, md)
            gi, gd = api.nn(tgt, q, md, ctx=ctx)
            assert np.array_equal(gi, oi) and np.array_equal(gd, od)
        s = q[:4000]
        o = orc.icp_align(s, tgt, 0.05, 8, dump_iteration=2)
        g = api.icp_align(s, tgt, 0.05, 8, dump_iteration=2, ctx=ctx)
        assert g["iterations"] == o["iterations"] and np.array_equal(g["corr_index"], o["corr_index"])


def test_nn_large_cloud_sort_paths(ctx):
    """> 327k points: the radix histograms outgrow the one-block scan (3-kernel scan path)."""
    rng = np.random.default_rng(22)
    tgt = rng.uniform(-1, 1, (420000, 3)).astype(np.float32)
    tgt[:, 2] = np.float32(0.3) * np.sin(3 * tgt[:, 0]) * np.cos(2 * tgt[:, 1])   # a surface, like a scan
    q = tgt[rng.choice(len(tgt), 50000)] + rng.normal(0, 0.004, (50000, 3)).astype(np.float32)
    oi, od = orc.KdTree(tgt).nn(q, 0.02)
    gi, gd = api.nn(tgt, q, 0.02, ctx=ctx)
    assert np.array_equal(gi, oi) and np.array_equal(gd, od)
    # percolation threshold of this sampling: 1225 clusters of 50..5666 points, many size ties
    lab, sz = api.euclidean_clusters(tgt[:350000], 0.004, 50, 350000, ctx=ctx)
    ol, osz = orc.euclidean_clusters(tgt[:350000], 0.004, 50, 350000)
    assert len(osz) > 1000 and np.array_equal(sz, osz) and np.array_equal(lab, ol)


# ------------------------------------------------------------------------------- ICP

@pytest.mark.parametrize("dump_it", [0, 4, 12])
def test_icp_p2p_matches_oracle(ctx, pair, dump_it):
    src, tgt = pair
    extent = float(np.ptp(tgt, axis=0).max())
    o = orc.icp_align(src, tgt, 0.02, 50, dump_iteration=dump_it, want_registered=True)
    g = api.icp_align(src, tgt, 0.02, 50, dump_iteration=dump_it, want_registered=True, ctx=ctx)
    assert g["iterations"] == o["iterations"] and g["state"] == o["state"] and g["converged"] == o["converged"]
    assert_transform_close(g["transformation"], o["transformation"], extent)
    assert abs(g["fitness"] - o["fitness"]) <= FIT_TOL * o["fitness"]
    if dump_it < o["iterations"]:
        assert np.array_equal(g["corr_index"], o["corr_index"])
        assert np.array_equal(g["corr_dist2"], o["corr_dist2"])
    assert np.abs(g["registered_xyz"] - o["registered_xyz"]).max() <= 1e-5 * extent


def test_icp_p2p_vs_pcl_like_float32_estimator(ctx, pair):
    """Against the oracle's float32 (PCL-faithful) Umeyama sums: north-star tolerances."""
    src, tgt = pair
    extent = float(np.ptp(tgt, axis=0).max())
    o = orc.icp_align(src, tgt, 0.02, 50, umeyama_f32=True)
    g = api.icp_align(src, tgt, 0.02, 50, ctx=ctx)
    assert g["iterations"] == o["iterations"]
    assert_transform_close(g["transformation"], o["transformation"], extent)
    # float32 estimator sums carry ~1e-5 relative noise of their own (sequential float32 sums
    # over ~3e4 terms), which shows up in the fitness at the few-1e-5 level
    assert abs(g["fitness"] - o["fitness"]) <= 1e-4 * o["fitness"]


def test_icp_p2plane_matches_oracle(ctx, pair_normals):
    src, tgt = pair_normals
    extent = float(np.ptp(tgt.xyz, axis=0).max())
    for dump_it in (0, 3):
        o = orc.icp_align(src, tgt, 0.02, 50, mode=1, dump_iteration=dump_it, want_registered=True)
        g = api.icp_align(src, tgt, 0.02, 50, mode=1, dump_iteration=dump_it, want_registered=True, ctx=ctx)
        assert g["iterations"] == o["iterations"] and g["state"] == o["state"]
        assert_transform_close(g["transformation"], o["transformation"], extent)
        assert abs(g["fitness"] - o["fitness"]) <= FIT_TOL * o["fitness"]
        if dump_it < o["iterations"]:
            assert np.array_equal(g["corr_index"], o["corr_index"])
        assert np.abs(g["registered_normal"] - o["registered_normal"]).max() <= 1e-5


def test_icp_recovers_known_motion(ctx):
    tgt = synth.kinect_view(0, scale=0.4, backdrop="panel", noise=False)
    T = synth.rigid(1.0, -2.0, 0.5, [0.004, -0.003, 0.002])
    src = synth.apply_transform(np.linalg.inv(T), tgt)
    g = api.icp_align(src, tgt, 0.05, 100, transformation_epsilon=1e-12, euclidean_fitness_epsilon=1e-9,
                      ctx=ctx)
    assert rot_angle(g["transformation"], T.astype(np.float32)) < 2e-4
    assert np.abs(g["transformation"][:3, 3] - T[:3, 3]).max() < 2e-4
    assert g["fitness"] < 1e-9


def test_icp_criteria_and_edge_cases(ctx, pair):
    src, tgt = pair
    # iteration cap
    o = orc.icp_align(src, tgt, 0.02, 3)
    g = api.icp_align(src, tgt, 0.02, 3, ctx=ctx)
    assert g["iterations"] == o["iterations"] == 3 and g["state"] == o["state"] == 1 and g["converged"]
    # not enough correspondences: far apart, tiny gate -> converged False (PCL semantics)
    far = src + np.float32(10.0)
    o = orc.icp_align(far, tgt, 0.001, 10)
    g = api.icp_align(far, tgt, 0.001, 10, ctx=ctx)
    assert g["state"] == o["state"] == 5 and not g["converged"] and g["iterations"] == o["iterations"] == 0
    assert np.array_equal(g["transformation"], np.eye(4, dtype=np.float32))
    # PCL-default gate (sqrt(DBL_MAX)): unbounded correspondences
    o = orc.icp_align(src[::7], tgt[::5], float(np.sqrt(np.finfo(np.float64).max)), 5)
    g = api.icp_align(src[::7], tgt[::5], float(np.sqrt(np.finfo(np.float64).max)), 5, ctx=ctx)
    assert g["iterations"] == o["iterations"]
    assert np.abs(g["transformation"] - o["transformation"]).max() < 1e-5
    # invalid arguments surface as errors, not crashes
    with pytest.raises(api.Lc3dError):
        api.icp_align(src, tgt, 0.02, 0, ctx=ctx)
    with pytest.raises(api.Lc3dError):
        api.icp_align(src, tgt, 0.02, 5, mode=1, ctx=ctx)  # p2plane without target normals


def test_icp_pcl_aos_layout_and_resident(ctx, pair_normals):
    """The 48-byte pcl::PointXYZRGBNormal layout and the resident (HBM) entry point give the
    same bits as packed host arrays."""
    src, tgt = pair_normals

    def aos(hc):
        a = np.zeros((hc.n, 12), dtype=np.float32)
        a[:, 0:3] = hc.xyz
        a[:, 3] = 1.0
        a[:, 4:7] = hc.normal
        a[:, 9] = hc.curvature
        return HostCloud.from_pcl_aos(a)

    ref = api.icp_align(src, tgt, 0.02, 30, mode=1, ctx=ctx)
    g = api.icp_align(aos(src), aos(tgt), 0.02, 30, mode=1, ctx=ctx)
    assert np.array_equal(g["transformation"], ref["transformation"]) and g["fitness"] == ref["fitness"]
    ds, dt = ctx.upload(src), ctx.upload(tgt)
    r = api.icp_align(ds, dt, 0.02, 30, mode=1, ctx=ctx)
    assert np.array_equal(r["transformation"], ref["transformation"]) and r["fitness"] == ref["fitness"]
    ds.free()
    dt.free()


def test_icp_packed_staging_of_pageable_aos_records(ctx, pair_normals, monkeypatch):
    """Pageable 48-byte pcl::PointXYZRGBNormal records above LC3D_PACK_MIN points are packed by the
    host thread pool into pinned staging chunks (12 bytes per point and field): same bits as the
    plain copy of the caller's records."""
    src, tgt = pair_normals

    def aos(hc):
        a = np.zeros((hc.n, 12), dtype=np.float32)
        a[:, 0:3] = hc.xyz
        a[:, 3] = 1.0
        a[:, 4:7] = hc.normal
        a[:, 9] = hc.curvature
        return HostCloud.from_pcl_aos(a)

    S, T = aos(src), aos(tgt)
    monkeypatch.setenv("LC3D_NO_PACK", "1")
    ref = api.icp_align(S, T, 0.02, 30, mode=1, want_registered=True, ctx=ctx)
    monkeypatch.delenv("LC3D_NO_PACK")
    monkeypatch.setenv("LC3D_PACK_MIN", "1000")
    for threads in ("1", "4"):
        monkeypatch.setenv("LC3D_PACK_THREADS", threads)  # (read when the context creates its pool)
        c2 = api.Context(0)
        g = api.icp_align(S, T, 0.02, 30, mode=1, want_registered=True, ctx=c2)
        c2.close()
        assert np.array_equal(g["transformation"], ref["transformation"]) and g["fitness"] == ref["fitness"]
        assert np.array_equal(g["registered_xyz"], ref["registered_xyz"])
        assert np.array_equal(g["registered_normal"], ref["registered_normal"])


def test_icp_deferred_target_normals_same_bits(ctx, pair_normals, monkeypatch):
    """Host-buffer point-to-plane ICP with iteration 0 searched before the target normals arrive
    (SEARCH_ONLY kernel + gather_normals_sorted + icp_estimate_kernel; chosen at run time from the
    measured PCIe rate) gives the same bits as the fused first iteration."""
    src, tgt = pair_normals
    out = {}
    for mode_env in ("0", "1"):
        monkeypatch.setenv("LC3D_DEFER_NORMALS", mode_env)
        out[mode_env] = api.icp_align(src, tgt, 0.02, 30, mode=1, dump_iteration=0, want_registered=True, ctx=ctx)
    a, b = out["0"], out["1"]
    assert a["iterations"] == b["iterations"] and a["state"] == b["state"] and a["fitness"] == b["fitness"]
    assert np.array_equal(a["transformation"], b["transformation"])
    assert np.array_equal(a["corr_index"], b["corr_index"]) and np.array_equal(a["corr_dist2"], b["corr_dist2"])
    assert np.array_equal(a["registered_xyz"], b["registered_xyz"])


def test_icp_pinned_and_pageable_host_buffers_agree(ctx, pair_normals):
    """Host-buffer ICP takes two upload paths: page-locked memory goes out with plain async copies
    of the caller's records, pageable memory is packed by the host thread pool into pinned staging
    chunks.  Same bits either way (and the same as with the deferred-normals iteration 0 off)."""
    src, tgt = pair_normals
    ref = api.icp_align(src, tgt, 0.02, 30, mode=1, want_registered=True, ctx=ctx)  # pageable numpy arrays
    arrs = [api.host_register(np.array(a, dtype=np.float32, order="C", copy=True))
            for a in (src.xyz, src.normal, tgt.xyz, tgt.normal)]
    try:
        g = api.icp_align(HostCloud(arrs[0], normal=arrs[1]), HostCloud(arrs[2], normal=arrs[3]), 0.02, 30, mode=1,
                          want_registered=True, ctx=ctx)
    finally:
        for a in arrs:
            api.host_unregister(a)
    assert np.array_equal(g["transformation"], ref["transformation"]) and g["fitness"] == ref["fitness"]
    assert g["iterations"] == ref["iterations"]
    assert np.array_equal(g["registered_xyz"], ref["registered_xyz"])
    assert np.array_equal(g["registered_normal"], ref["registered_normal"])


def test_icp_equivariance_property(ctx, pair):
    """Size-independent property: moving both clouds by a rigid motion M
    conjugates the solution, T' = M T M^-1, up to float rounding."""
    src, tgt = pair
    M = synth.rigid(10, 20, -5, [0.3, -0.2, 0.1])
    g0 = api.icp_align(src, tgt, 0.02, 50, ctx=ctx)
    g1 = api.icp_align(synth.apply_transform(M, src), synth.apply_transform(M, tgt), 0.02, 50, ctx=ctx)
    Texp = M @ g0["transformation"].astype(np.float64) @ np.linalg.inv(M)
    assert rot_angle(g1["transformation"], Texp.astype(np.float32)) < 2e-3
    assert abs(g1["fitness"] - g0["fitness"]) < 0.05 * g0["fitness"]


# ------------------------------------------------------------------------------- kNN

@pytest.mark.parametrize("k", [1, 8, 31, 51, 101])
def test_knn_bit_exact(ctx, pair, k):
    _, tgt = pair
    oi, od = orc.KdTree(tgt).knn(tgt, k)
    gi, gd = api.knn(tgt, k, ctx=ctx)
    assert np.array_equal(gd, od)
    assert np.array_equal(gi, oi)


def test_knn_external_queries_and_small_clouds(ctx):
    rng = np.random.default_rng(5)
    c = rng.normal(0, 1, (20000, 3)).astype(np.float32)
    q = rng.normal(0, 2, (3000, 3)).astype(np.float32)
    oi, od = orc.KdTree(c).knn(q, 16)
    gi, gd = api.knn(c, 16, queries=q, ctx=ctx)
    assert np.array_equal(gi, oi) and np.array_equal(gd, od)
    # fewer points than k: missing neighbours are (-1, inf)
    small = rng.normal(0, 1, (7, 3)).astype(np.float32)
    oi, od = orc.KdTree(small).knn(small, 10)
    gi, gd = api.knn(small, 10, ctx=ctx)
    assert np.array_equal(gi, oi) and np.array_equal(gd, od)
    # many duplicates (ties broken by index) — more coincident points than the buffer holds
    dup = np.repeat(rng.normal(0, 1, (40, 3)).astype(np.float32), 200, axis=0)
    oi, od = orc.KdTree(dup).knn(dup[::50], 20)
    gi, gd = api.knn(dup, 20, queries=dup[::50], ctx=ctx)
    assert np.array_equal(gi, oi) and np.array_equal(gd, od)


def test_knn_sparse_outliers_and_density_jumps(ctx):
    """The block-growing restart of the k-NN search: isolated outliers metres away from a dense
    surface (what StatisticalOutlierRemoval exists for), a 100x density jump, and a line of points
    (degenerate extent in two axes)."""
    rng = np.random.default_rng(11)
    dense = np.c_[rng.uniform(0, 0.2, (30000, 2)), rng.normal(0, 1e-4, 30000)]
    sparse = np.c_[rng.uniform(0.2, 2.2, (3000, 2)), rng.normal(0, 1e-3, 3000)]
    outliers = rng.uniform(-5, 5, (60, 3))
    cloud = np.concatenate([dense, sparse, outliers]).astype(np.float32)
    for k in (5, 51):
        oi, od = orc.KdTree(cloud).knn(cloud, k)
        gi, gd = api.knn(cloud, k, ctx=ctx)
        assert np.array_equal(gd, od) and np.array_equal(gi, oi)
    kept, md, st = api.sor(cloud, 50, 1.0, ctx=ctx)
    okept, omd, ost = orc.sor(cloud, 50, 1.0)
    assert np.array_equal(kept, okept) and np.array_equal(md, omd)
    line = np.c_[np.linspace(0, 1, 5000), np.zeros(5000), np.zeros(5000)].astype(np.float32)
    oi, od = orc.KdTree(line).knn(line, 12)
    gi, gd = api.knn(line, 12, ctx=ctx)
    assert np.array_equal(gd, od) and np.array_equal(gi, oi)


# ----------------------------------------------------------------------------- normals

def test_normals_match_oracle(ctx, pair):
    _, tgt = pair
    for k, vp in ((30, (0.0, 0.0, 0.0)), (50, tuple(orc.centroid(tgt)[:3]))):
        on, oc = orc.normals(tgt, k, vp)
        gn, gc = api.normals(tgt, k, vp, ctx=ctx)
        # angle from the cross product (arccos of a float32 dot is useless below ~5e-4 rad)
        ang = np.arcsin(np.clip(np.linalg.norm(np.cross(on.astype(np.float64), gn.astype(np.float64)), axis=1), 0, 1))
        ang = np.where(np.sum(on.astype(np.float64) * gn, axis=1) < 0, np.pi - ang, ang)
        # identical neighbour sets and float32 op order; only atan2f/cosf/sinf differ between
        # glibc and CUDA (<= 2 ulp) -> sub-microradian differences
        assert np.nanmax(ang) < 2e-3
        assert np.nanmedian(ang) < 1e-6
        assert (ang < 1e-5).mean() > 0.995
        assert np.nanmax(np.abs(gc - oc)) < 1e-3
        assert np.isnan(gn).sum() == np.isnan(on).sum()


def test_centroid_bit_exact(ctx, pair):
    _, tgt = pair
    assert np.array_equal(api.centroid(tgt, ctx=ctx), orc.centroid(tgt))


# --------------------------------------------------------------------------------- SOR

@pytest.mark.parametrize("k,mul", [(50, 1.0), (50, 5.0), (8, 0.5)])
def test_sor_mask_bit_exact(ctx, pair, k, mul):
    _, tgt = pair
    rng = np.random.default_rng(k)
    cloud = tgt.copy()
    out = rng.choice(len(cloud), 300, replace=False)
    cloud[out] += rng.normal(0, 0.03, (300, 3)).astype(np.float32)  # injected outliers
    ok, omd, ost = orc.sor(cloud, k, mul)
    gk, gmd, gst = api.sor(cloud, k, mul, ctx=ctx)
    assert np.array_equal(gmd, omd)          # per-point mean distances, float32 bits
    assert np.array_equal(gk, ok)            # kept-index list == mask, order preserved
    assert np.allclose(gst, ost, rtol=1e-12)
    okn, _, _ = orc.sor(cloud, k, mul, negative=True)
    gkn, _, _ = api.sor(cloud, k, mul, negative=True, ctx=ctx)
    assert np.array_equal(gkn, okn)
    assert len(gk) + len(gkn) == len(cloud)


# --------------------------------------------------------------------------- VoxelGrid

@pytest.mark.parametrize("leaf", [0.002, 0.01, 0.05, 1.0])  # 1.0 = the CLI default: a handful of huge voxels
def test_voxel_grid_bit_exact(ctx, pair_normals, leaf):
    _, tgt = pair_normals
    rng = np.random.default_rng(1)
    rgba = rng.integers(0, 2**32, tgt.n, dtype=np.uint32)
    cloud = HostCloud(tgt.xyz, normal=tgt.normal, rgba=rgba, curvature=tgt.curvature)
    o = orc.voxel_grid(cloud, leaf)
    g = api.voxel_grid(cloud, leaf, ctx=ctx)
    assert np.array_equal(g["voxel_of_point"], o["voxel_of_point"])  # occupancy, assignment, order
    assert np.array_equal(g["xyz"], o["xyz"])                        # centroids: same summation order
    assert np.array_equal(g["rgba"], o["rgba"])
    assert np.array_equal(g["curvature"], o["curvature"])
    assert np.nanmax(np.abs(g["normal"] - o["normal"])) <= 1e-6


def test_voxel_grid_overflow_guard_and_edges(ctx, pair):
    _, tgt = pair
    # leaf so small that dx*dy*dz > INT32_MAX: PCL returns the input unfiltered
    o = orc.voxel_grid(tgt, 1e-5)
    g = api.voxel_grid(tgt, 1e-5, ctx=ctx)
    assert o["overflow"] and len(g["xyz"]) == len(tgt) and np.array_equal(g["xyz"], tgt)
    # one voxel swallowing everything (all coordinates positive -> a single voxel)
    shifted = tgt[:5000] + np.float32(10.0)
    g = api.voxel_grid(shifted, 100.0, ctx=ctx)
    o = orc.voxel_grid(shifted, 100.0)
    assert len(g["xyz"]) == 1 and np.array_equal(g["xyz"], o["xyz"])
    g = api.voxel_grid(tgt[:5000], 100.0, ctx=ctx)
    o = orc.voxel_grid(tgt[:5000], 100.0)
    assert np.array_equal(g["xyz"], o["xyz"]) and np.array_equal(g["voxel_of_point"], o["voxel_of_point"])
    # idempotence-style property: filtering the centroids with the same leaf keeps their count
    g1 = api.voxel_grid(tgt, 0.01, ctx=ctx)
    g2 = api.voxel_grid(g1["xyz"], 0.01, ctx=ctx)
    assert len(g2["xyz"]) <= len(g1["xyz"])


# --------------------------------------------------------------------------- transform

def test_transform_bit_exact(ctx, pair_normals):
    src, _ = pair_normals
    T = synth.rigid(3, -4, 5, [0.1, 0.2, -0.3]).astype(np.float32)
    ox, on = orc.transform(src, T)
    gx, gn = api.transform(src, T, ctx=ctx)
    assert np.array_equal(gx, ox) and np.array_equal(gn, on)


def test_empty_and_tiny_clouds(ctx):
    e = np.zeros((0, 3), np.float32)
    one = np.array([[1.0, 2.0, 3.0]], np.float32)
    few = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]], np.float32)
    for s, t in ((e, few), (few, e), (one, one), (e, e)):
        g = api.icp_align(s, t, 0.5, 5, ctx=ctx)
        o = orc.icp_align(s, t, 0.5, 5)
        assert g["state"] == o["state"] == 5 and not g["converged"] and g["iterations"] == 0
        assert np.array_equal(g["transformation"], np.eye(4, dtype=np.float32))
    g = api.icp_align(few + np.float32(0.01), few, 0.5, 20, want_registered=True, ctx=ctx)
    o = orc.icp_align(few + np.float32(0.01), few, 0.5, 20, want_registered=True)
    assert g["iterations"] == o["iterations"] and np.abs(g["transformation"] - o["transformation"]).max() < 1e-5
    assert api.voxel_grid(e, 0.1, ctx=ctx)["xyz"].shape == (0, 3)
    assert len(api.sor(e, 5, 1.0, ctx=ctx)[0]) == 0
    idx, d2 = api.nn(e, few, ctx=ctx)
    assert idx.tolist() == [-1] * 4
    nrm, curv = api.normals(few[:2], 5, ctx=ctx)
    assert np.isnan(nrm).all()


@pytest.mark.parametrize("radius", [0.002, 0.01, 0.05])
def test_box_dedup_bit_exact(ctx, pair, radius):
    """accumulate_clouds.cpp:100-111 de-dup: kept-index list identical to the O(N*M) restatement."""
    src, tgt = pair
    rng = np.random.default_rng(2)
    s = src[rng.choice(len(src), 4000, replace=False)].copy()
    s[:50] = tgt[:50] + np.float32(radius)            # on the inclusive box face
    s[50] = np.nan
    t = tgt[rng.choice(len(tgt), 6000, replace=False)]
    assert np.array_equal(api.box_dedup(s, t, radius, ctx=ctx), orc.box_dedup(s, t, radius))
    assert len(api.box_dedup(s, np.zeros((0, 3), np.float32), radius, ctx=ctx)) == len(s) - 1   # only the NaN goes
    assert len(api.box_dedup(tgt, tgt, radius, ctx=ctx)) == 0                                    # everything is redundant


@pytest.mark.parametrize("tol,lo", [(0.004, 30), (0.02, 100), (0.0015, 1)])
def test_euclidean_clusters_bit_exact(ctx, pair, tol, lo):
    """cluster_extraction.cpp:94-101: per-point cluster ranks identical to the BFS restatement."""
    src, tgt = pair
    rng = np.random.default_rng(5)
    pts = np.concatenate([tgt[rng.choice(len(tgt), 15000, replace=False)],
                          rng.uniform(-0.6, 0.6, (300, 3)).astype(np.float32) + np.array([0, 0, 0.9], np.float32)])
    pts[123] = np.nan
    g_lab, g_sz = api.euclidean_clusters(pts, tol, lo, len(pts), ctx=ctx)
    o_lab, o_sz = orc.euclidean_clusters(pts, tol, lo, len(pts))
    assert np.array_equal(g_sz, o_sz) and np.array_equal(g_lab, o_lab)
    assert g_lab[123] == -1
    # PCL-named wrapper: index lists, largest cluster first, indices ascending
    ec = api.EuclideanClusterExtraction(ctx)
    ec.setInputCloud(pts)
    ec.setClusterTolerance(tol)
    ec.setMinClusterSize(lo)
    ec.setMaxClusterSize(len(pts))
    cl = ec.extract()
    assert [len(c) for c in cl] == o_sz.tolist()
    for r, c in enumerate(cl):
        assert np.array_equal(c, np.flatnonzero(o_lab == r))


def test_euclidean_clusters_edge_cases(ctx):
    e = np.zeros((0, 3), np.float32)
    assert api.euclidean_clusters(e, 0.1, 1, 10, ctx=ctx)[1].tolist() == []
    two = np.array([[0, 0, 0], [0.5, 0, 0]], np.float32)
    assert api.euclidean_clusters(two, 0.5, 1, 2, ctx=ctx)[1].tolist() == [1, 1]        # strict radius
    assert api.euclidean_clusters(two, 0.5000001, 1, 2, ctx=ctx)[1].tolist() == [2]
    lab, sz = api.euclidean_clusters(two, 0.6, 3, 10, ctx=ctx)                            # too small: no cluster
    assert sz.tolist() == [] and lab.tolist() == [-1, -1]
    same = np.zeros((100, 3), np.float32)                                                # duplicates
    lab, sz = api.euclidean_clusters(same, 1e-3, 1, 100, ctx=ctx)
    assert sz.tolist() == [100] and (lab == 0).all()


# ------------------------------------------------------------------- ABI argument validation

def test_abi_rejects_bad_layouts_with_a_message(ctx):
    """Strides that are not multiples of 4 bytes would fault on the device (4-byte loads): they are
    rejected up front with LC3D_ERR_INVALID and a message; so are non-positive k / leaf sizes."""
    import ctypes as C
    from lowcost3dreconstruction_b200 import _capi
    lib = _capi.load()
    raw = np.zeros(15 * 64 + 16, dtype=np.uint8)  # packed 15-byte xyz+rgb records
    c = _capi.Cloud()
    c.n, c.xyz, c.xyz_stride = 64, raw.ctypes.data, 15
    idx = np.empty(64, np.int32)
    d2 = np.empty(64, np.float32)
    rc = lib.lc3d_nn(ctx._h, C.byref(c), None, 0.0, idx.ctypes.data, d2.ctypes.data)
    assert rc == -1 and b"multiple of 4" in lib.lc3d_last_error(ctx._h)
    pts = np.random.default_rng(0).uniform(-1, 1, (200, 3)).astype(np.float32)
    with pytest.raises(api.Lc3dError, match="mean_k"):
        api.sor(pts, 0, 1.0, ctx=ctx)
    with pytest.raises(api.Lc3dError, match="neighbours must be positive"):
        api.normals(pts, 0, ctx=ctx)
    with pytest.raises(api.Lc3dError, match="leaf size"):
        api.voxel_grid(pts, 0.0, ctx=ctx)
    # the context stays usable after a rejected call (no sticky device error)
    i2, _ = api.nn(pts, ctx=ctx)
    assert np.array_equal(i2, np.arange(200))
    # a resident source without normals returns no registered normals; mixing resident / host raises
    ds, dt = ctx.upload(pts), ctx.upload(pts)
    r = api.icp_align(ds, dt, 0.05, 5, want_registered=True, ctx=ctx)
    assert "registered_normal" not in r and r["registered_xyz"].shape == (200, 3)
    with pytest.raises(api.Lc3dError, match="both be resident"):
        api.icp_align(ds, pts, 0.05, 5, ctx=ctx)
