"""Generates tests/golden/golden_independent.npz WITHOUT the oracle: an independent numpy / scipy
implementation of the PCL semantics of SURVEY.md Appendix A (run from the repo root:
`python tests/golden/make_golden_independent.py`).

Nothing here imports oracle/ or the CUDA library: nearest neighbours come from
scipy.spatial.cKDTree (candidates re-ranked with PCL's float32 ((dx^2)+dy^2)+dz^2 and the lower
index on exact ties, the convention include/lc3d.h documents), the estimators from numpy.linalg
(SVD / solve) in float64, the incremental float32 transformCloud from plain float32 numpy
arithmetic (numpy never fuses multiply-add), DefaultConvergenceCriteria and getFitnessScore from
their definitions (A.3, A.4), VoxelGrid / StatisticalOutlierRemoval / NormalEstimation likewise
(A.6-A.8, normals in float64: compared under an angular tolerance).  The fixtures pin BOTH our
CPU oracle (tests/test_oracle.py) and the CUDA path (tests/test_gpu_golden.py) against a second
implementation; the reference itself ships no vectors and cannot be built here.
"""
import os
import sys

import numpy as np
from scipy.spatial import cKDTree

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from lowcost3dreconstruction_b200 import synth  # noqa: E402  (data generation only)

f32 = np.float32


def d2_f32(q, p):
    """PCL/FLANN squared distance: float32, ((dx*dx) + dy*dy) + dz*dz, no FMA."""
    d = (q.astype(f32) - p.astype(f32)).astype(f32)
    return ((d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]).astype(f32) + d[..., 2] * d[..., 2]).astype(f32)


def nn_exact(tree, tgt, q, kcand=4):
    """index and float32 d2 of the nearest neighbour under the float32 metric (ties -> lower index)."""
    _, cand = tree.query(q.astype(np.float64), k=kcand)
    d2 = d2_f32(q[:, None, :], tgt[cand])
    order = np.lexsort((cand, d2), axis=1)[:, 0]
    rows = np.arange(len(q))
    best = cand[rows, order]
    bd2 = d2[rows, order]
    # safety: the k-th candidate must be strictly farther than the winner, otherwise widen
    unsafe = d2.max(axis=1) <= bd2
    assert not unsafe.any(), "increase kcand"
    return best.astype(np.int32), bd2


def transform_f32(T, X):
    """transformCloud: ((T0*x + T1*y) + T2*z) + T3 per row, float32."""
    T = T.astype(f32)
    x, y, z = X[:, 0], X[:, 1], X[:, 2]
    out = np.empty_like(X)
    for r in range(3):
        out[:, r] = (((T[r, 0] * x + T[r, 1] * y).astype(f32) + T[r, 2] * z).astype(f32) + T[r, 3]).astype(f32)
    return out


def estimate_p2p(s, d):
    """Umeyama without scale (A.5), float64."""
    s, d = s.astype(np.float64), d.astype(np.float64)
    ms, md = s.mean(0), d.mean(0)
    S = (d - md).T @ (s - ms) / len(s)
    U, _, Vt = np.linalg.svd(S)
    D = np.eye(3)
    if np.linalg.det(U) * np.linalg.det(Vt) < 0:
        D[2, 2] = -1
    R = U @ D @ Vt
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = md - R @ ms
    return T


def estimate_p2plane(s, d, n):
    """TransformationEstimationPointToPlaneLLS (A.5): float32 products widened to float64."""
    s, d, n = s.astype(f32), d.astype(f32), n.astype(f32)
    a = (n[:, 2] * s[:, 1] - n[:, 1] * s[:, 2]).astype(f32)
    b = (n[:, 0] * s[:, 2] - n[:, 2] * s[:, 0]).astype(f32)
    c = (n[:, 1] * s[:, 0] - n[:, 0] * s[:, 1]).astype(f32)
    J = np.stack([a, b, c, n[:, 0], n[:, 1], n[:, 2]], axis=1).astype(np.float64)
    # r = n.d - n.s evaluated in float32 like the reference expression
    r = (n[:, 0] * d[:, 0] + n[:, 1] * d[:, 1] + n[:, 2] * d[:, 2] - n[:, 0] * s[:, 0] - n[:, 1] * s[:, 1]
         - n[:, 2] * s[:, 2]).astype(np.float64)
    x = np.linalg.solve(J.T @ J, J.T @ r)
    al, be, ga = x[:3]
    sa, ca, sb, cb, sg, cg = np.sin(al), np.cos(al), np.sin(be), np.cos(be), np.sin(ga), np.cos(ga)
    T = np.eye(4)
    T[0, :3] = [cg * cb, -sg * ca + cg * sb * sa, sg * sa + cg * sb * ca]
    T[1, :3] = [sg * cb, cg * ca + sg * sb * sa, -cg * sa + sg * sb * ca]
    T[2, :3] = [-sb, cb * sa, cb * ca]
    T[:3, 3] = x[3:]
    return T


def icp(src, tgt, nrm, mode, max_dist=0.02, max_iter=50, teps=1e-9, feps=1e-3, dump=(0, 3)):
    tree = cKDTree(tgt.astype(np.float64))
    X = src.astype(f32).copy()
    final = np.eye(4, dtype=f32)
    prev_mse = np.finfo(np.float64).max
    gate = max_dist * max_dist
    dumps = {}
    it, state = 0, 0
    while True:
        j, d2 = nn_exact(tree, tgt, X)
        has = d2.astype(np.float64) <= gate
        if it in dump:
            dumps[it] = np.where(has, j, -1).astype(np.int32)
        if has.sum() < 3:
            state = 5
            break
        if mode == 0:
            T = estimate_p2p(X[has], tgt[j[has]])
        else:
            T = estimate_p2plane(X[has], tgt[j[has]], nrm[j[has]])
        T = T.astype(f32)
        X = transform_f32(T, X)
        final = (T @ final).astype(f32)  # Matrix4f product (float32 accumulate order is immaterial at 1e-7)
        it += 1
        mse = d2[has].astype(np.float64).sum() / has.sum()
        if it >= max_iter:
            state = 1
            break
        cos_angle = 0.5 * float(f32(f32(f32(T[0, 0] + T[1, 1]) + T[2, 2]) - f32(1)))
        tsq = float(f32(T[0, 3] * T[0, 3])) + float(f32(T[1, 3] * T[1, 3])) + float(f32(T[2, 3] * T[2, 3]))
        if cos_angle >= 1.0 - teps and tsq <= teps:
            state = 2
            break
        if abs(mse - prev_mse) < 1e-12:
            state = 3
            break
        if abs(mse - prev_mse) / prev_mse < feps:
            state = 4
            break
        prev_mse = mse
    reg = transform_f32(final, src.astype(f32))
    _, fd2 = nn_exact(tree, tgt, reg)
    return dict(T=final, iterations=it, state=state, fitness=fd2.astype(np.float64).mean(), dumps=dumps)


def normals_pca(tgt, k):
    tree = cKDTree(tgt.astype(np.float64))
    _, idx = tree.query(tgt.astype(np.float64), k=k)
    P = tgt[idx].astype(np.float64)
    C = P - P.mean(1, keepdims=True)
    cov = np.einsum("nki,nkj->nij", C, C) / k
    w, v = np.linalg.eigh(cov)
    n = v[:, :, 0]
    flip = np.einsum("ni,ni->n", n, -tgt.astype(np.float64)) < 0
    n[flip] *= -1
    curv = np.abs(w[:, 0] / np.maximum(w.sum(1), 1e-300))
    return n.astype(f32), curv.astype(f32), idx


def sor(tgt, k, mul):
    tree = cKDTree(tgt.astype(np.float64))
    _, idx = tree.query(tgt.astype(np.float64), k=k + 1)
    d2 = d2_f32(tgt[:, None, :], tgt[idx])
    d2.sort(axis=1)
    mean = (np.sqrt(d2[:, 1:].astype(np.float64)).sum(1) / k).astype(f32)
    s = mean.astype(np.float64).sum()
    sq = (mean * mean).astype(f32).astype(np.float64).sum()
    n = len(tgt)
    mu = s / n
    var = (sq - s * s / n) / (n - 1)
    thr = mu + mul * np.sqrt(var)
    return np.nonzero(mean.astype(np.float64) <= thr)[0].astype(np.int32), mean, thr


def voxel(tgt, leaf):
    inv = f32(1.0) / f32(leaf)
    mn = tgt.min(0)
    minb = np.floor((mn * inv).astype(f32)).astype(np.int64)
    ijk = (np.floor((tgt * inv).astype(f32)) - minb.astype(f32)).astype(np.int64)
    mx = tgt.max(0)
    div = np.floor((mx * inv).astype(f32)).astype(np.int64) - minb + 1
    key = ijk[:, 0] + ijk[:, 1] * div[0] + ijk[:, 2] * div[0] * div[1]
    uniq, inverse = np.unique(key, return_inverse=True)
    cent = np.zeros((len(uniq), 3))
    np.add.at(cent, inverse, tgt.astype(np.float64))
    cent /= np.bincount(inverse)[:, None]
    return inverse.astype(np.int32), cent.astype(f32)


def main():
    tgt = synth.kinect_view(0, scale=0.2, backdrop="panel")
    src = synth.kinect_view(1, scale=0.2, backdrop="panel")
    out = dict(src=src, tgt=tgt)
    tree = cKDTree(tgt.astype(np.float64))
    j, d2 = nn_exact(tree, tgt, src)
    out["nn_idx"] = np.where(d2.astype(np.float64) <= 0.02 * 0.02, j, -1).astype(np.int32)
    out["nn_d2"] = d2
    nrm, curv, _ = normals_pca(tgt, 20)
    out["normals"], out["curvature"] = nrm, curv
    for name, mode in (("p2p", 0), ("p2plane", 1)):
        r = icp(src, tgt, nrm, mode)
        out[f"icp_{name}_T"] = r["T"]
        out[f"icp_{name}_meta"] = np.array([r["iterations"], r["state"]], dtype=np.int64)
        out[f"icp_{name}_fitness"] = np.array([r["fitness"]])
        for it, c in r["dumps"].items():
            out[f"icp_{name}_corr{it}"] = c
        print(name, "iterations", r["iterations"], "state", r["state"], "fitness", r["fitness"])
    # criteria exits the turntable pair never takes: TRANSFORM (a loose epsilon) and ABS_MSE (a copy
    # of the target perturbed at the 1e-7 level)
    r = icp(src, tgt, nrm, 1, teps=1e-5, feps=0.0)
    out["icp_transform_exit_meta"] = np.array([r["iterations"], r["state"]], dtype=np.int64)
    out["icp_transform_exit_T"] = r["T"]
    rng = np.random.default_rng(11)
    near = (tgt.astype(np.float64) + rng.normal(0, 1e-7, tgt.shape)).astype(f32)
    r = icp(near, tgt, nrm, 0, teps=0.0, feps=0.0)
    out["near"] = near
    out["icp_abs_mse_exit_meta"] = np.array([r["iterations"], r["state"]], dtype=np.int64)
    print("transform-exit", out["icp_transform_exit_meta"], "abs-mse-exit", out["icp_abs_mse_exit_meta"])
    kept, mean, thr = sor(tgt, 10, 1.0)
    out["sor_kept"], out["sor_mean"], out["sor_thr"] = kept, mean, np.array([thr])
    vop, cent = voxel(tgt, 0.02)
    out["vox_of_point"], out["vox_xyz"] = vop, cent
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_independent.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
