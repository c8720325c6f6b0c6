"""CPU tests of the oracle (the checker itself): against independent implementations
(scipy cKDTree, numpy SVD / eigh, brute force, analytic ground truth) and the committed golden
fixtures.  The reference has no tests or golden vectors of its own and PCL
cannot be built here, so this is how the restatement is pinned.  No GPU needed."""
import os

import numpy as np
import pytest
from scipy.spatial import cKDTree

from lowcost3dreconstruction_b200 import synth
from lowcost3dreconstruction_b200._capi import HostCloud
from oracle import oracle as orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_small.npz")


@pytest.fixture(scope="module")
def pair():
    return synth.kinect_view(1, scale=0.2, backdrop="panel"), synth.kinect_view(0, scale=0.2, backdrop="panel")


def float_d2(a, b):
    d = a.astype(np.float32) - b.astype(np.float32)
    r = d[:, 0] * d[:, 0]
    r = r + d[:, 1] * d[:, 1]
    r = r + d[:, 2] * d[:, 2]
    return r.astype(np.float32)


def test_nn_matches_scipy_and_float_arithmetic(pair):
    src, tgt = pair
    idx, d2 = orc.KdTree(tgt).nn(src)
    dd, ii = cKDTree(tgt).query(src)
    # float32 ((dx^2)+dy^2)+dz^2 of the returned match, bit for bit
    assert np.array_equal(d2, float_d2(src, tgt[idx]))
    # same neighbour as scipy unless the two candidates tie in float32
    diff = idx != ii
    assert np.array_equal(float_d2(src[diff], tgt[ii[diff]]) >= d2[diff], np.ones(diff.sum(), bool))
    assert np.allclose(np.sqrt(d2), dd, rtol=1e-5, atol=1e-7)


def test_nn_brute_force_with_ties_and_gate():
    rng = np.random.default_rng(0)
    tgt = rng.integers(0, 6, (400, 3)).astype(np.float32)  # lattice: many exact ties + duplicates
    q = rng.integers(0, 6, (300, 3)).astype(np.float32) + np.float32(0.5)
    idx, d2 = orc.KdTree(tgt).nn(q, 1.0)
    for i in range(len(q)):
        dd = float_d2(np.repeat(q[i:i + 1], len(tgt), 0), tgt)
        j = int(np.flatnonzero(dd == dd.min())[0])  # lowest index among ties
        if float(dd[j]) > 1.0:
            assert idx[i] == -1
        else:
            assert idx[i] == j and d2[i] == dd[j]


def test_knn_matches_scipy(pair):
    _, tgt = pair
    k = 17
    idx, d2 = orc.KdTree(tgt).knn(tgt[:3000], k)
    dd, ii = cKDTree(tgt).query(tgt[:3000], k=k)
    assert np.all(np.diff(d2, axis=1) >= 0)               # ascending
    assert np.array_equal(idx[:, 0], np.arange(3000))     # self first (distance 0, lowest index)
    assert np.allclose(np.sqrt(d2), dd, rtol=1e-5, atol=1e-7)
    same = (idx == ii).mean()
    assert same > 0.98                                     # differences only at float32 ties
    for r in np.flatnonzero((idx != ii).any(axis=1))[:50]:
        assert np.array_equal(np.sort(d2[r]), np.sort(float_d2(np.repeat(tgt[r:r + 1], k, 0), tgt[ii[r]])))


def test_kabsch_rotation_against_numpy_svd():
    rng = np.random.default_rng(1)
    for trial in range(200):
        S = rng.normal(size=(3, 3))
        if trial % 5 == 0:
            S[:, 2] = S[:, 0] * 0.3 - S[:, 1]  # rank 2
        U, _, Vt = np.linalg.svd(S)
        D = np.diag([1.0, 1.0, np.sign(np.linalg.det(U) * np.linalg.det(Vt))])
        Rn = U @ D @ Vt
        R = orc.kabsch_rotation(S)
        assert np.allclose(R @ R.T, np.eye(3), atol=1e-10) and np.linalg.det(R) > 0
        if abs(np.linalg.svd(S, compute_uv=False)[1:].min()) > 1e-6 or trial % 5 == 0:
            assert np.trace(R @ S.T) >= np.trace(Rn @ S.T) - 1e-9  # maximises tr(R S^T)


def test_eigen33_against_numpy_eigh():
    rng = np.random.default_rng(2)
    for _ in range(300):
        A = rng.normal(size=(3, 3))
        C = (A @ A.T).astype(np.float32)
        ev, v = orc.eigen33(C)
        w, V = np.linalg.eigh(C.astype(np.float64))
        assert abs(ev - w[0]) <= 2e-5 * w[2] + 1e-7
        if (w[1] - w[0]) > 1e-2 * w[2]:
            assert abs(abs(float(v @ V[:, 0])) - 1.0) < 1e-3


def test_icp_recovers_known_motion():
    tgt = synth.kinect_view(0, scale=0.2, backdrop="panel", noise=False)
    T = synth.rigid(1.0, -2.0, 0.5, [0.004, -0.003, 0.002])
    src = synth.apply_transform(np.linalg.inv(T), tgt)
    for f32 in (False, True):
        r = orc.icp_align(src, tgt, 0.05, 100, transformation_epsilon=1e-12, euclidean_fitness_epsilon=1e-9,
                          umeyama_f32=f32)
        assert np.abs(r["transformation"] - T).max() < 3e-4
        assert r["fitness"] < 1e-8
    nrm, curv = orc.normals(tgt, 15)
    r = orc.icp_align(src, HostCloud(tgt, normal=nrm), 0.05, 100, mode=1, transformation_epsilon=1e-12,
                      euclidean_fitness_epsilon=1e-9)
    assert np.abs(r["transformation"] - T).max() < 3e-4


def test_icp_convergence_states(pair):
    src, tgt = pair
    r = orc.icp_align(src, tgt, 0.02, 2)
    assert r["iterations"] == 2 and r["state"] == 1 and r["converged"]          # ITERATIONS
    r = orc.icp_align(src, tgt, 0.02, 50)
    assert r["state"] == 4 and r["converged"] and r["iterations"] < 50          # REL_MSE (1e-3)
    r = orc.icp_align(src + np.float32(5.0), tgt, 0.001, 10)
    assert r["state"] == 5 and not r["converged"] and r["iterations"] == 0      # NO_CORRESPONDENCES
    r = orc.icp_align(tgt, tgt, 0.02, 50)
    assert r["converged"] and r["iterations"] <= 2 and r["fitness"] == 0.0
    assert np.allclose(r["transformation"], np.eye(4), atol=1e-6)


def test_normals_against_numpy_pca(pair):
    _, tgt = pair
    k = 20
    nrm, curv = orc.normals(tgt, k)
    _, ii = cKDTree(tgt).query(tgt, k=k)
    sub = np.arange(0, len(tgt), 7)
    P = tgt[ii[sub]].astype(np.float64)
    C = np.einsum("nki,nkj->nij", P - P.mean(1, keepdims=True), P - P.mean(1, keepdims=True)) / k
    w, V = np.linalg.eigh(C)
    ref = V[:, :, 0]
    cosang = np.abs(np.sum(ref * nrm[sub], axis=1))
    # PCL's float32 single-pass covariance is noisy: degrees, not microradians
    assert np.median(np.degrees(np.arccos(np.clip(cosang, -1, 1)))) < 1.5
    assert (np.degrees(np.arccos(np.clip(cosang, -1, 1))) < 10).mean() > 0.97
    # flipped toward the viewpoint (origin), unit length, curvature in [0, 1/3]
    assert np.all(np.sum(-tgt * nrm, axis=1) >= -1e-6)
    assert np.allclose(np.linalg.norm(nrm, axis=1), 1.0, atol=1e-5)
    assert np.all((curv >= 0) & (curv <= 1.0 / 3.0 + 1e-3))


def test_sor_against_numpy(pair):
    _, tgt = pair
    k, mul = 12, 1.5
    kept, md, st = orc.sor(tgt, k, mul)
    dd, _ = cKDTree(tgt).query(tgt, k=k + 1)
    ref = dd[:, 1:].sum(1) / k
    assert np.allclose(md, ref, rtol=2e-5)
    mean, sd = ref.mean(), ref.std(ddof=1)
    assert abs(st[0] - mean) < 1e-6 * mean and abs(st[1] - sd) < 1e-4 * sd
    mask = np.zeros(len(tgt), bool)
    mask[kept] = True
    edge = np.abs(ref - st[2]) < 1e-6 * st[2]
    assert np.array_equal(mask[~edge], (ref <= st[2])[~edge])
    assert np.all(np.diff(kept) > 0)  # input order preserved
    neg, _, _ = orc.sor(tgt, k, mul, negative=True)
    assert len(neg) + len(kept) == len(tgt) and not np.intersect1d(neg, kept).size


def test_voxel_grid_against_numpy(pair):
    _, tgt = pair
    leaf = 0.01
    v = orc.voxel_grid(tgt, leaf)
    inv = np.float32(1.0) / np.float32(leaf)
    ijk = np.floor(tgt * inv).astype(np.int64)
    ijk -= ijk.min(0)
    dims = ijk.max(0) + 1
    lin = ijk[:, 0] + dims[0] * (ijk[:, 1] + dims[1] * ijk[:, 2])
    uniq, inverse = np.unique(lin, return_inverse=True)
    assert len(v["xyz"]) == len(uniq)                          # occupancy
    assert np.array_equal(v["voxel_of_point"], inverse)        # assignment + ascending-index output order
    cen = np.zeros((len(uniq), 3))
    np.add.at(cen, inverse, tgt.astype(np.float64))
    cen /= np.bincount(inverse)[:, None]
    assert np.abs(v["xyz"] - cen).max() < 1e-6
    assert orc.voxel_grid(tgt, 1e-5)["overflow"]               # PCL's index-overflow guard


def test_transform_and_centroid(pair):
    _, tgt = pair
    T = synth.rigid(3, -4, 5, [0.1, 0.2, -0.3])
    x, _ = orc.transform(tgt, T)
    assert np.abs(x - synth.apply_transform(T, tgt)).max() < 1e-6
    assert np.abs(orc.centroid(tgt)[:3] - tgt.astype(np.float64).mean(0)).max() < 1e-4


def test_empty_and_tiny_inputs():
    e = np.zeros((0, 3), np.float32)
    one = np.array([[1.0, 2.0, 3.0]], np.float32)
    idx, d2 = orc.KdTree(e).nn(one)
    assert idx[0] == -1 and np.isinf(d2[0])
    idx, d2 = orc.KdTree(one).knn(one, 3)
    assert idx.tolist() == [[0, -1, -1]]
    r = orc.icp_align(one, one, 0.1, 5)
    assert r["state"] == 5 and not r["converged"]
    assert orc.voxel_grid(e, 0.1)["xyz"].shape == (0, 3)
    nrm, curv = orc.normals(np.array([[0, 0, 0], [1, 0, 0]], np.float32), 5)
    assert np.isnan(nrm).all() and np.isnan(curv).all()  # fewer than 3 neighbours -> NaN (PCL)


def test_oracle_reproduces_golden_fixtures():
    g = np.load(GOLDEN)
    src, tgt = g["src"], g["tgt"]
    kt = orc.KdTree(tgt)
    idx, d2 = kt.nn(src, 0.02)
    assert np.array_equal(idx, g["nn_idx"]) and np.array_equal(d2, g["nn_d2"])
    ki, kd = kt.knn(tgt, 12)
    assert np.array_equal(ki, g["knn_idx"]) and np.array_equal(kd, g["knn_d2"])
    nrm, curv = orc.normals(tgt, 12)
    assert np.allclose(nrm, g["normals"], atol=1e-6, equal_nan=True)
    T = HostCloud(tgt, normal=g["normals"], curvature=g["curvature"])
    for name, mode in (("p2p", 0), ("p2plane", 1)):
        r = orc.icp_align(src, T, 0.02, 50, mode=mode, dump_iteration=1)
        assert [r["iterations"], r["state"], int(r["converged"])] == g[f"icp_{name}_meta"].tolist()
        assert np.abs(r["transformation"] - g[f"icp_{name}_T"]).max() < 1e-6
        assert np.array_equal(r["corr_index"], g[f"icp_{name}_corr1"])
        assert abs(r["fitness"] - g[f"icp_{name}_fitness"][0]) < 1e-9
    kept, md, st = orc.sor(tgt, 10, 1.0)
    assert np.array_equal(kept, g["sor_kept"]) and np.array_equal(md, g["sor_mean"])
    v = orc.voxel_grid(T, 0.02)
    assert np.array_equal(v["voxel_of_point"], g["vox_of_point"]) and np.array_equal(v["xyz"], g["vox_xyz"])


def test_box_dedup_against_numpy():
    rng = np.random.default_rng(4)
    tgt = rng.uniform(0, 1, (600, 3)).astype(np.float32)
    src = rng.uniform(-0.2, 1.2, (900, 3)).astype(np.float32)
    src[10] = np.nan
    src[20:25] = tgt[5:10] + np.float32(0.03)       # exactly on a box face for r = 0.03 (inclusive bound)
    for r in (0.03, 0.1, 0.0):
        kept = orc.box_dedup(src, tgt, r)
        lo = (tgt.astype(np.float64) - r).astype(np.float32)
        hi = (tgt.astype(np.float64) + r).astype(np.float32)
        inside = ((src[:, None, :] >= lo[None]) & (src[:, None, :] <= hi[None])).all(-1).any(-1)
        want = np.flatnonzero(~inside & np.isfinite(src).all(1))
        assert np.array_equal(kept, want)


def test_euclidean_clusters_against_connected_components():
    """extractEuclideanClusters == connected components of the strict d2 < (float)(tol^2) graph,
    filtered by size, ranked by size descending (ties: lower first index)."""
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components
    rng = np.random.default_rng(7)
    a = (rng.normal(size=(900, 3)) * 0.05).astype(np.float32)
    b = (rng.normal(size=(600, 3)) * 0.05).astype(np.float32) + np.array([1, 0, 0], np.float32)
    c = rng.uniform(-3, 3, size=(150, 3)).astype(np.float32)
    pts = np.concatenate([a, b, c])
    rng.shuffle(pts)
    pts[17] = np.nan
    tol = 0.05
    for lo, hi in ((50, len(pts)), (1, len(pts)), (2, 700)):
        labels, sizes = orc.euclidean_clusters(pts, tol, lo, hi)
        fin = np.isfinite(pts).all(1)
        d = (pts[:, None, :] - pts[None, :, :]).astype(np.float32) ** 2
        d2 = (d[:, :, 0] + d[:, :, 1]) + d[:, :, 2]
        adj = (d2 < np.float32(tol * tol)) & fin[:, None] & fin[None, :]
        nc, comp = connected_components(coo_matrix(adj), directed=False)
        cnt = np.bincount(comp, minlength=nc)
        first = np.full(nc, len(pts))
        np.minimum.at(first, comp, np.arange(len(pts)))
        keep = [k for k in range(nc) if lo <= cnt[k] <= hi and fin[first[k]]]
        keep.sort(key=lambda k: (-cnt[k], first[k]))
        want = np.full(len(pts), -1, np.int32)
        for r, k in enumerate(keep):
            want[comp == k] = r
        assert np.array_equal(labels, want)
        assert np.array_equal(sizes, [cnt[k] for k in keep])
    # the radius test is strict: two points exactly `tol` apart are NOT neighbours
    two = np.array([[0, 0, 0], [0.5, 0, 0]], np.float32)
    assert orc.euclidean_clusters(two, 0.5, 1, 2)[1].tolist() == [1, 1]
    assert orc.euclidean_clusters(two, 0.5000001, 1, 2)[1].tolist() == [2]


def test_single_iteration_estimators_against_numpy(pair):
    """One ICP iteration of the oracle vs an independent numpy restatement of the published
    estimators: Umeyama without scale, and the point-to-plane LLS
    (J = [n x s-ish cross terms, n], r = n.(d - s), x = (J^T J)^-1 J^T r, R = Rz Ry Rx)."""
    src, tgt = pair
    nrm, _ = orc.normals(tgt, 15)
    ok = np.isfinite(nrm).all(1)
    tgt, nrm = tgt[ok], nrm[ok]
    max_d = 0.02
    tree = orc.KdTree(tgt)
    idx, d2 = tree.nn(src, max_d)
    m = idx >= 0
    s = src[m].astype(np.float64)
    d = tgt[idx[m]].astype(np.float64)
    n = nrm[idx[m]].astype(np.float64)
    # --- point-to-point: Umeyama / Kabsch
    cnt, T = tree.one_iteration(src, tgt, max_d, 0)
    assert cnt == int(m.sum())
    mu_s, mu_d = s.mean(0), d.mean(0)
    S = (d - mu_d).T @ (s - mu_s) / len(s)
    U, _, Vt = np.linalg.svd(S)
    D = np.diag([1.0, 1.0, np.sign(np.linalg.det(U) * np.linalg.det(Vt))])
    R = U @ D @ Vt
    t = mu_d - R @ mu_s
    assert np.abs(T[:3, :3] - R).max() < 2e-6 and np.abs(T[:3, 3] - t).max() < 2e-6
    # --- point-to-plane LLS
    cnt, T = tree.one_iteration(src, HostCloud(tgt, normal=nrm), max_d, 1)
    assert cnt == int(m.sum())
    J = np.concatenate([np.cross(s, n), n], axis=1)
    r = np.einsum("ij,ij->i", n, d - s)
    x = np.linalg.solve(J.T @ J, J.T @ r)
    al, be, ga = x[:3]
    Rx = np.array([[1, 0, 0], [0, np.cos(al), -np.sin(al)], [0, np.sin(al), np.cos(al)]])
    Ry = np.array([[np.cos(be), 0, np.sin(be)], [0, 1, 0], [-np.sin(be), 0, np.cos(be)]])
    Rz = np.array([[np.cos(ga), -np.sin(ga), 0], [np.sin(ga), np.cos(ga), 0], [0, 0, 1]])
    assert np.abs(T[:3, :3] - Rz @ Ry @ Rx).max() < 2e-6 and np.abs(T[:3, 3] - x[3:]).max() < 2e-6
    assert np.array_equal(T[3], [0, 0, 0, 1])


def test_convergence_criteria_replayed_from_the_iteration_log(pair):
    """DefaultConvergenceCriteria replayed in numpy on the oracle's per-iteration
    log (correspondence count, mse): the loop must stop at the FIRST iteration whose relative
    mse change drops below euclidean_fitness_epsilon, and not before."""
    src, tgt = pair
    for eps in (1e-3, 1e-4, 1e-2):
        # transformation_epsilon tiny: the TRANSFORM exit (|t|^2 <= eps) must not pre-empt the mse test
        r = orc.icp_align(src, tgt, 0.02, 50, euclidean_fitness_epsilon=eps, transformation_epsilon=1e-30)
        mse = r["log"][:, 1]
        assert r["state"] == 4 and len(mse) == r["iterations"]
        rel = np.abs(np.diff(mse)) / mse[:-1]
        stop = np.flatnonzero(rel < eps)
        assert len(stop) and stop[0] + 2 == r["iterations"]       # iteration k+1 sees |mse_k+1 - mse_k| / mse_k
        assert (np.abs(np.diff(mse))[: stop[0]] >= 1e-12).all()    # the absolute test never fired earlier
        assert r["last_mse"] == mse[-1] and r["last_correspondences"] == r["log"][-1, 0]
    # mse is the mean SQUARED kd-tree distance of the correspondences found before the update
    idx, d2 = orc.KdTree(tgt).nn(src, 0.02)
    r = orc.icp_align(src, tgt, 0.02, 1)
    assert r["last_correspondences"] == (idx >= 0).sum()
    assert abs(r["last_mse"] - d2[idx >= 0].astype(np.float64).mean()) <= 1e-12 * r["last_mse"]


def test_fitness_score_is_the_unbounded_mean_nn_distance(pair):
    """getFitnessScore: every source point counts, no max-distance gate; the
    registered cloud is ONE application of the composed transform to the original source."""
    src, tgt = pair
    r = orc.icp_align(src, tgt, 0.02, 50, want_registered=True)
    reg = r["registered_xyz"]
    xyz, _ = orc.transform(src, r["transformation"])
    assert np.array_equal(reg, xyz)
    d2 = cKDTree(tgt.astype(np.float64)).query(reg.astype(np.float64))[0] ** 2
    assert abs(r["fitness"] - d2.mean()) <= 1e-6 * d2.mean()
    assert (d2 > 0.02 ** 2).any()          # the sum really includes points beyond the gate


# ------------------------------------------------------------------ independent golden fixture

def test_oracle_matches_independent_golden():
    """tests/golden/golden_independent.npz comes from tests/golden/make_golden_independent.py — a
    numpy / scipy implementation that shares no code with the oracle.  Index sets bit-exact,
    float32 distances bit-exact, transforms / fitness within the north-star tolerances, all five
    convergence exits of DefaultConvergenceCriteria that the fixtures reach."""
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_independent.npz"))
    src, tgt = g["src"], g["tgt"]
    oi, od = orc.KdTree(tgt).nn(src, 0.02)
    assert np.array_equal(oi, g["nn_idx"])
    assert np.array_equal(od[oi >= 0], g["nn_d2"][oi >= 0])
    T = HostCloud(tgt, normal=g["normals"], curvature=g["curvature"])
    extent = float(np.ptp(tgt, axis=0).max())
    for name, mode in (("p2p", 0), ("p2plane", 1)):
        for it in (0, 3):
            r = orc.icp_align(src, T, 0.02, 50, mode=mode, dump_iteration=it)
            assert np.array_equal(r["corr_index"], g[f"icp_{name}_corr{it}"]), (name, it)
        assert [r["iterations"], r["state"]] == g[f"icp_{name}_meta"].tolist()
        assert np.abs(r["transformation"] - g[f"icp_{name}_T"])[:3, :3].max() < 1e-5
        assert np.abs(r["transformation"] - g[f"icp_{name}_T"])[:3, 3].max() < 1e-5 * extent
        assert abs(r["fitness"] - g[f"icp_{name}_fitness"][0]) <= 1e-5 * r["fitness"]
    r = orc.icp_align(src, T, 0.02, 50, transformation_epsilon=1e-5, euclidean_fitness_epsilon=0.0, mode=1)
    assert [r["iterations"], r["state"]] == g["icp_transform_exit_meta"].tolist() and r["state"] == 2
    r = orc.icp_align(g["near"], tgt, 0.02, 50, transformation_epsilon=0.0, euclidean_fitness_epsilon=0.0, mode=0)
    assert [r["iterations"], r["state"]] == g["icp_abs_mse_exit_meta"].tolist() and r["state"] == 3
    kept, mean, st = orc.sor(tgt, 10, 1.0)
    assert np.array_equal(kept, g["sor_kept"]) and np.array_equal(mean, g["sor_mean"])
    v = orc.voxel_grid(tgt, 0.02)
    assert np.array_equal(v["voxel_of_point"], g["vox_of_point"]) and np.abs(v["xyz"] - g["vox_xyz"]).max() < 1e-6
