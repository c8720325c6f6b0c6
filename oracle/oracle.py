"""ctypes wrapper of the CPU oracle (oracle/lc3d_oracle.cpp).  TEST INFRASTRUCTURE ONLY.

May be imported only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs — never from lowcost3dreconstruction_b200/.  PARITY UNPINNED:
the reference ships no golden vectors and PCL is absent (see lc3d_oracle.cpp header).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from lowcost3dreconstruction_b200._capi import Cloud, HostCloud, IcpOutputs, IcpParams, IcpResult

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liblc3d_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "lc3d_oracle.cpp")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return LIB_PATH


def load():
    global _lib
    if _lib is not None:
        return _lib
    build()
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    cp = C.POINTER(Cloud)
    lib.orc_kdtree_build.argtypes = [cp]
    lib.orc_kdtree_build.restype = vp
    lib.orc_kdtree_free.argtypes = [vp]
    lib.orc_kdtree_free.restype = None
    lib.orc_kdtree_knn.argtypes = [vp, cp, i32, vp, vp]
    lib.orc_kdtree_knn.restype = None
    lib.orc_kdtree_nn.argtypes = [vp, cp, f64, vp, vp]
    lib.orc_kdtree_nn.restype = None
    lib.orc_icp_align.argtypes = [cp, cp, C.POINTER(IcpParams), i32, C.POINTER(IcpResult),
                                  C.POINTER(IcpOutputs), vp]
    lib.orc_icp_align.restype = C.c_int
    lib.orc_icp_one_iteration.argtypes = [vp, cp, cp, f64, i32, C.POINTER(C.c_float)]
    lib.orc_icp_one_iteration.restype = i64
    lib.orc_centroid.argtypes = [cp, C.POINTER(C.c_float)]
    lib.orc_centroid.restype = None
    lib.orc_normals.argtypes = [cp, i32, C.POINTER(C.c_float), vp, vp, vp]
    lib.orc_normals.restype = C.c_int
    lib.orc_sor.argtypes = [cp, i32, f64, i32, vp, C.POINTER(i64), vp, C.POINTER(f64)]
    lib.orc_sor.restype = C.c_int
    lib.orc_voxel_grid.argtypes = [cp, C.POINTER(C.c_float), vp, vp, vp, vp, vp, C.POINTER(i64)]
    lib.orc_voxel_grid.restype = C.c_int
    lib.orc_box_dedup.argtypes = [cp, cp, f64, vp, C.POINTER(i64)]
    lib.orc_box_dedup.restype = C.c_int
    lib.orc_euclidean_clusters.argtypes = [cp, f64, i64, i64, vp, vp, i64, C.POINTER(i64)]
    lib.orc_euclidean_clusters.restype = C.c_int
    lib.orc_transform.argtypes = [cp, C.POINTER(C.c_float), vp, vp]
    lib.orc_transform.restype = None
    lib.orc_kabsch_rotation.argtypes = [C.POINTER(f64), C.POINTER(f64)]
    lib.orc_kabsch_rotation.restype = None
    lib.orc_eigen33.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float)]
    lib.orc_eigen33.restype = None
    _lib = lib
    return lib


def _hc(x) -> HostCloud:
    return x if isinstance(x, HostCloud) else HostCloud(x)


class KdTree:
    def __init__(self, cloud):
        self.cloud = _hc(cloud)
        self._h = load().orc_kdtree_build(self.cloud.ref())

    def __del__(self):
        if getattr(self, "_h", None):
            load().orc_kdtree_free(self._h)
            self._h = None

    def knn(self, queries, k: int):
        q = _hc(queries)
        idx = np.empty((q.n, k), dtype=np.int32)
        d2 = np.empty((q.n, k), dtype=np.float32)
        load().orc_kdtree_knn(self._h, q.ref(), k, idx.ctypes.data, d2.ctypes.data)
        return idx, d2

    def nn(self, queries, max_dist: float = 0.0):
        q = _hc(queries)
        idx = np.empty(q.n, dtype=np.int32)
        d2 = np.empty(q.n, dtype=np.float32)
        load().orc_kdtree_nn(self._h, q.ref(), float(max_dist), idx.ctypes.data, d2.ctypes.data)
        return idx, d2

    def one_iteration(self, src, tgt, max_dist: float, mode: int):
        T = (C.c_float * 16)()
        n = load().orc_icp_one_iteration(self._h, _hc(src).ref(), _hc(tgt).ref(), float(max_dist), mode, T)
        return int(n), np.array(T, dtype=np.float32).reshape(4, 4)


def icp_align(src, tgt, max_correspondence_distance=0.1, max_iterations=50, transformation_epsilon=1e-9,
              euclidean_fitness_epsilon=1e-3, mode=0, compute_fitness=True, dump_iteration=-1,
              umeyama_f32=False, want_registered=False):
    s, t = _hc(src), _hc(tgt)
    p = IcpParams(max_correspondence_distance, transformation_epsilon, euclidean_fitness_epsilon,
                  max_iterations, mode, int(compute_fitness), dump_iteration)
    r = IcpResult()
    o = IcpOutputs()
    out = {}
    if dump_iteration >= 0:
        out["corr_index"] = np.empty(s.n, dtype=np.int32)
        out["corr_dist2"] = np.empty(s.n, dtype=np.float32)
        o.corr_index = out["corr_index"].ctypes.data
        o.corr_dist2 = out["corr_dist2"].ctypes.data
    if want_registered:
        out["registered_xyz"] = np.empty((s.n, 3), dtype=np.float32)
        o.registered_xyz = out["registered_xyz"].ctypes.data
        if s.normal is not None:
            out["registered_normal"] = np.empty((s.n, 3), dtype=np.float32)
            o.registered_normal = out["registered_normal"].ctypes.data
    log = np.zeros((max_iterations + 1, 2), dtype=np.float64)
    load().orc_icp_align(s.ref(), t.ref(), C.byref(p), int(umeyama_f32), C.byref(r), C.byref(o),
                         log.ctypes.data)
    out.update(
        transformation=np.array(r.transformation, dtype=np.float32).reshape(4, 4),
        fitness=r.fitness, converged=bool(r.converged), iterations=r.iterations, state=r.state,
        last_mse=r.last_mse, last_correspondences=r.last_correspondences,
        log=log[: max(r.iterations, 1)],
    )
    return out


def centroid(cloud):
    out = (C.c_float * 4)()
    load().orc_centroid(_hc(cloud).ref(), out)
    return np.array(out, dtype=np.float32)


def normals(cloud, k: int, viewpoint=(0.0, 0.0, 0.0), knn_idx=None):
    c = _hc(cloud)
    vp = (C.c_float * 3)(*viewpoint)
    nrm = np.empty((c.n, 3), dtype=np.float32)
    curv = np.empty(c.n, dtype=np.float32)
    ki = None
    if knn_idx is not None:
        ki = np.ascontiguousarray(knn_idx, dtype=np.int32)
        assert ki.shape == (c.n, k)
    load().orc_normals(c.ref(), k, vp, None if ki is None else ki.ctypes.data, nrm.ctypes.data,
                       curv.ctypes.data)
    return nrm, curv


def sor(cloud, mean_k: int, stddev_mul: float, negative: bool = False):
    c = _hc(cloud)
    kept = np.empty(c.n, dtype=np.int32)
    cnt = C.c_int64(0)
    md = np.empty(c.n, dtype=np.float32)
    stats = (C.c_double * 3)()
    load().orc_sor(c.ref(), mean_k, float(stddev_mul), int(negative), kept.ctypes.data, C.byref(cnt),
                   md.ctypes.data, stats)
    return kept[: cnt.value].copy(), md, np.array(stats)


def voxel_grid(cloud, leaf):
    c = _hc(cloud)
    lf = (C.c_float * 3)(*([leaf] * 3 if np.isscalar(leaf) else leaf))
    n = c.n
    xyz = np.empty((n, 3), dtype=np.float32)
    nrm = np.empty((n, 3), dtype=np.float32) if c.normal is not None else None
    rgba = np.empty(n, dtype=np.uint32) if c.rgba is not None else None
    curv = np.empty(n, dtype=np.float32) if c.curvature is not None else None
    vox = np.empty(n, dtype=np.int32)
    cnt = C.c_int64(0)
    rc = load().orc_voxel_grid(c.ref(), lf, xyz.ctypes.data, None if nrm is None else nrm.ctypes.data,
                               None if rgba is None else rgba.ctypes.data,
                               None if curv is None else curv.ctypes.data, vox.ctypes.data, C.byref(cnt))
    m = cnt.value
    return dict(xyz=xyz[:m].copy(), normal=None if nrm is None else nrm[:m].copy(),
                rgba=None if rgba is None else rgba[:m].copy(),
                curvature=None if curv is None else curv[:m].copy(), voxel_of_point=vox, overflow=rc == 1)


def box_dedup(src, tgt, radius: float):
    s, t = _hc(src), _hc(tgt)
    kept = np.empty(max(s.n, 1), dtype=np.int32)
    cnt = C.c_int64(0)
    load().orc_box_dedup(s.ref(), t.ref(), float(radius), kept.ctypes.data, C.byref(cnt))
    return kept[: cnt.value].copy()


def euclidean_clusters(cloud, tolerance: float, min_size: int, max_size: int):
    c = _hc(cloud)
    labels = np.empty(max(c.n, 1), dtype=np.int32)
    sizes = np.zeros(max(c.n, 1), dtype=np.int64)
    cnt = C.c_int64(0)
    load().orc_euclidean_clusters(c.ref(), float(tolerance), int(min_size), int(max_size), labels.ctypes.data,
                                  sizes.ctypes.data, sizes.size, C.byref(cnt))
    return labels[: c.n].copy(), sizes[: cnt.value].copy()


def transform(cloud, T):
    c = _hc(cloud)
    Tm = (C.c_float * 16)(*np.asarray(T, dtype=np.float32).reshape(16))
    xyz = np.empty((c.n, 3), dtype=np.float32)
    nrm = np.empty((c.n, 3), dtype=np.float32) if c.normal is not None else None
    load().orc_transform(c.ref(), Tm, xyz.ctypes.data, None if nrm is None else nrm.ctypes.data)
    return xyz, nrm


def kabsch_rotation(S):
    S = np.ascontiguousarray(S, dtype=np.float64).reshape(9)
    R = np.empty(9, dtype=np.float64)
    load().orc_kabsch_rotation(S.ctypes.data_as(C.POINTER(C.c_double)), R.ctypes.data_as(C.POINTER(C.c_double)))
    return R.reshape(3, 3)


def eigen33(Cm):
    Cm = np.ascontiguousarray(Cm, dtype=np.float32).reshape(9)
    ev = C.c_float(0)
    v = (C.c_float * 3)()
    load().orc_eigen33(Cm.ctypes.data_as(C.POINTER(C.c_float)), C.byref(ev), v)
    return float(ev.value), np.array(v, dtype=np.float32)
