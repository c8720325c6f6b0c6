"""Host-side mirror of the PCL interface the reference's hot-path tools call, over the C ABI.

Function level (`icp_align`, `knn`, `normals`, `sor`, `voxel_grid`, `transform`) and PCL-named
classes (`IterativeClosestPoint`, `NormalEstimation`, `VoxelGrid`,
`StatisticalOutlierRemoval`) with the setters used at
pcl_tools/fine_registration.cpp:105-126, normal_estimation.cpp:84-108,
cloud_downsampling.cpp:73-76 and outlier_removal.cpp:80-93.  Everything computes on the GPU
through liblc3d.so; there is no CPU path here.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi
from ._capi import HostCloud, IcpOutputs, IcpParams, IcpResult, POINT_TO_PLANE, POINT_TO_POINT


class Lc3dError(RuntimeError):
    pass


class Context:
    """One device + one stream (struct lc3d_ctx).  Not thread-safe; use one per thread/GPU."""

    def __init__(self, device: int = 0, stream: int | None = None):
        self._lib = _capi.load()
        h = C.c_void_p()
        rc = self._lib.lc3d_create(int(device), C.c_void_p(stream or 0), C.byref(h))
        if rc != 0:
            raise Lc3dError(f"lc3d_create failed ({rc}): {self._lib.lc3d_last_error(None).decode()}")
        self._h = h
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            self._lib.lc3d_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc: int, what: str):
        if rc != 0:
            raise Lc3dError(f"{what} failed ({rc}): {self._lib.lc3d_last_error(self._h).decode()}")

    @property
    def launch_count(self) -> int:
        return int(self._lib.lc3d_launch_count(self._h))

    def grid_info(self) -> dict:
        out = (C.c_double * 8)()
        self._lib.lc3d_debug_grid_info(self._h, out)
        return dict(cell=out[0], dims=(int(out[1]), int(out[2]), int(out[3])), cells=int(out[4]), points=int(out[5]))

    # ---- resident clouds -------------------------------------------------------------
    def upload(self, cloud) -> "DeviceCloud":
        hc = _hc(cloud)
        h = C.c_void_p()
        self._check(self._lib.lc3d_cloud_upload(self._h, hc.ref(), C.byref(h)), "lc3d_cloud_upload")
        return DeviceCloud(self, h, hc.n, has_normal=hc.normal is not None)


class DeviceCloud:
    """A cloud resident in HBM (struct lc3d_dcloud): xyz, optionally normals + curvature."""

    def __init__(self, ctx: Context, handle, n: int, has_normal: bool = False):
        self.ctx, self._h, self.n, self.has_normal = ctx, handle, n, has_normal

    def download(self):
        """(xyz (n,3), normal (n,3) or None, curvature (n,) or None) as host arrays."""
        xyz = np.empty((self.n, 3), dtype=np.float32)
        nrm = np.empty((self.n, 3), dtype=np.float32) if self.has_normal else None
        curv = np.empty(self.n, dtype=np.float32) if self.has_normal else None
        self.ctx._check(self.ctx._lib.lc3d_cloud_download(
            self.ctx._h, self._h, xyz.ctypes.data, None if nrm is None else nrm.ctypes.data,
            None if curv is None else curv.ctypes.data), "lc3d_cloud_download")
        return xyz, nrm, curv

    def free(self):
        if self._h and self.ctx._h:
            self.ctx._lib.lc3d_cloud_free(self.ctx._h, self._h)
        self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def host_register(a: np.ndarray) -> np.ndarray:
    """Page-locks the array's buffer in place (lc3d_host_register = cudaHostRegister): uploads from
    it are direct DMA instead of staged pageable copies.  Returns the (contiguous float32) array;
    pair with host_unregister before the buffer is freed."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    rc = _capi.load().lc3d_host_register(C.c_void_p(a.ctypes.data), a.nbytes)
    if rc != 0:
        raise Lc3dError(f"lc3d_host_register failed ({rc})")
    return a


def host_unregister(a: np.ndarray) -> None:
    _capi.load().lc3d_host_unregister(C.c_void_p(a.ctypes.data))


_default_ctx: Context | None = None


def default_context() -> Context:
    global _default_ctx
    if _default_ctx is None or _default_ctx._h is None:
        _default_ctx = Context(0)
    return _default_ctx


def _hc(x) -> HostCloud:
    return x if isinstance(x, HostCloud) else HostCloud(x)


def _result_dict(r: IcpResult) -> dict:
    return dict(
        transformation=np.array(r.transformation, dtype=np.float32).reshape(4, 4),
        fitness=r.fitness, converged=bool(r.converged), iterations=int(r.iterations), state=int(r.state),
        last_mse=r.last_mse, last_correspondences=int(r.last_correspondences),
        ms=dict(upload=r.ms_upload, index=r.ms_index, loop=r.ms_loop, fitness=r.ms_fitness,
                download=r.ms_download, total=r.ms_total),
    )


def icp_align(src, tgt, max_correspondence_distance=0.1, max_iterations=50, transformation_epsilon=1e-9,
              euclidean_fitness_epsilon=1e-3, mode=POINT_TO_POINT, compute_fitness=True, dump_iteration=-1,
              want_registered=False, ctx: Context | None = None, registered_out=None) -> dict:
    """pcl::IterativeClosestPoint align + getFinalTransformation + hasConverged +
    getFitnessScore (fine_registration.cpp:105-126).  src/tgt: arrays (n,3), HostCloud, or
    DeviceCloud (both resident).  registered_out: optional (xyz, normal) float32 (n,3) arrays to
    receive the registered cloud (e.g. pinned memory: the download is then a true async DMA)."""
    ctx = ctx or default_context()
    p = IcpParams(float(max_correspondence_distance), float(transformation_epsilon),
                  float(euclidean_fitness_epsilon), int(max_iterations), int(mode), int(bool(compute_fitness)),
                  int(dump_iteration))
    r = IcpResult()
    o = IcpOutputs()
    out = {}
    resident = isinstance(src, DeviceCloud)
    if resident != isinstance(tgt, DeviceCloud):
        raise Lc3dError("icp_align: source and target must both be resident (DeviceCloud) or both host clouds")
    n = src.n if resident else _hc(src).n
    if not resident:
        src, tgt = _hc(src), _hc(tgt)
    src_has_normal = src.has_normal if resident else src.normal is not None
    if dump_iteration >= 0:
        out["corr_index"] = np.empty(n, dtype=np.int32)
        out["corr_dist2"] = np.empty(n, dtype=np.float32)
        o.corr_index = out["corr_index"].ctypes.data
        o.corr_dist2 = out["corr_dist2"].ctypes.data
    if registered_out is not None:
        rx, rn = registered_out
        assert rx.dtype == np.float32 and rx.shape == (n, 3) and rx.flags.c_contiguous
        out["registered_xyz"] = rx
        o.registered_xyz = rx.ctypes.data
        if rn is not None and src_has_normal:
            assert rn.dtype == np.float32 and rn.shape == (n, 3) and rn.flags.c_contiguous
            out["registered_normal"] = rn
            o.registered_normal = rn.ctypes.data
    elif want_registered:
        out["registered_xyz"] = np.empty((n, 3), dtype=np.float32)
        o.registered_xyz = out["registered_xyz"].ctypes.data
        if src_has_normal:
            out["registered_normal"] = np.empty((n, 3), dtype=np.float32)
            o.registered_normal = out["registered_normal"].ctypes.data
    if resident:
        rc = ctx._lib.lc3d_icp_align_resident(ctx._h, src._h, tgt._h, C.byref(p), C.byref(r), C.byref(o))
    else:
        rc = ctx._lib.lc3d_icp_align(ctx._h, src.ref(), tgt.ref(), C.byref(p), C.byref(r), C.byref(o))
    ctx._check(rc, "lc3d_icp_align")
    out.update(_result_dict(r))
    return out


def prepare_view(cloud, leaf_size: float = 0.0, sor_mean_k: int = 0, sor_stddev_mul: float = 1.0, normals_k: int = 0,
                 viewpoint=(0.0, 0.0, 0.0), ctx: Context | None = None):
    """VoxelGrid -> StatisticalOutlierRemoval -> NormalEstimation chained on the device
    (lc3d_prepare_view): returns (DeviceCloud, (n_after_voxel, n_after_sor, n_final)).  A stage
    whose parameter is <= 0 is skipped."""
    ctx = ctx or default_context()
    c = _hc(cloud)
    p = _capi.PrepareParams(float(leaf_size), int(sor_mean_k), float(sor_stddev_mul), int(normals_k),
                            (C.c_float * 3)(*[float(v) for v in viewpoint]))
    h = C.c_void_p()
    counts = (C.c_int64 * 3)()
    ctx._check(ctx._lib.lc3d_prepare_view(ctx._h, c.ref(), C.byref(p), C.byref(h), counts), "lc3d_prepare_view")
    return DeviceCloud(ctx, h, int(counts[2]), has_normal=normals_k > 0 and counts[2] > 0), tuple(int(x) for x in counts)


def nn(cloud, queries=None, max_dist: float = 0.0, ctx: Context | None = None):
    ctx = ctx or default_context()
    c = _hc(cloud)
    q = c if queries is None else _hc(queries)
    idx = np.empty(q.n, dtype=np.int32)
    d2 = np.empty(q.n, dtype=np.float32)
    ctx._check(ctx._lib.lc3d_nn(ctx._h, c.ref(), None if queries is None else q.ref(), float(max_dist),
                                idx.ctypes.data, d2.ctypes.data), "lc3d_nn")
    return idx, d2


def knn(cloud, k: int, queries=None, ctx: Context | None = None):
    ctx = ctx or default_context()
    c = _hc(cloud)
    q = c if queries is None else _hc(queries)
    idx = np.empty((q.n, k), dtype=np.int32)
    d2 = np.empty((q.n, k), dtype=np.float32)
    ctx._check(ctx._lib.lc3d_knn(ctx._h, c.ref(), None if queries is None else q.ref(), int(k),
                                 idx.ctypes.data, d2.ctypes.data), "lc3d_knn")
    return idx, d2


def centroid(cloud, ctx: Context | None = None) -> np.ndarray:
    ctx = ctx or default_context()
    out = (C.c_float * 4)()
    ctx._check(ctx._lib.lc3d_centroid(ctx._h, _hc(cloud).ref(), out), "lc3d_centroid")
    return np.array(out, dtype=np.float32)


def normals(cloud, k: int, viewpoint=(0.0, 0.0, 0.0), ctx: Context | None = None):
    ctx = ctx or default_context()
    c = _hc(cloud)
    vp = (C.c_float * 3)(*[float(v) for v in viewpoint])
    nrm = np.empty((c.n, 3), dtype=np.float32)
    curv = np.empty(c.n, dtype=np.float32)
    ctx._check(ctx._lib.lc3d_normals(ctx._h, c.ref(), int(k), vp, nrm.ctypes.data, curv.ctypes.data),
               "lc3d_normals")
    return nrm, curv


def sor(cloud, mean_k: int, stddev_mul: float, negative: bool = False, ctx: Context | None = None):
    ctx = ctx or default_context()
    c = _hc(cloud)
    kept = np.empty(max(c.n, 1), dtype=np.int32)
    cnt = C.c_int64(0)
    md = np.empty(max(c.n, 1), dtype=np.float32)
    stats = (C.c_double * 3)()
    ctx._check(ctx._lib.lc3d_sor(ctx._h, c.ref(), int(mean_k), float(stddev_mul), int(bool(negative)),
                                 kept.ctypes.data, C.byref(cnt), md.ctypes.data, stats), "lc3d_sor")
    return kept[: cnt.value].copy(), md[: c.n], np.array(stats)


def voxel_grid(cloud, leaf, ctx: Context | None = None) -> dict:
    ctx = ctx or default_context()
    c = _hc(cloud)
    lf = (C.c_float * 3)(*([float(leaf)] * 3 if np.isscalar(leaf) else [float(v) for v in leaf]))
    n = max(c.n, 1)
    xyz = np.empty((n, 3), dtype=np.float32)
    nrm = np.empty((n, 3), dtype=np.float32) if c.normal is not None else None
    rgba = np.empty(n, dtype=np.uint32) if c.rgba is not None else None
    curv = np.empty(n, dtype=np.float32) if c.curvature is not None else None
    vox = np.empty(n, dtype=np.int32)
    cnt = C.c_int64(0)
    ctx._check(ctx._lib.lc3d_voxel_grid(
        ctx._h, c.ref(), lf, xyz.ctypes.data, None if nrm is None else nrm.ctypes.data,
        None if rgba is None else rgba.ctypes.data, None if curv is None else curv.ctypes.data,
        vox.ctypes.data, C.byref(cnt)), "lc3d_voxel_grid")
    m = cnt.value
    return dict(xyz=xyz[:m].copy(), normal=None if nrm is None else nrm[:m].copy(),
                rgba=None if rgba is None else rgba[:m].copy(),
                curvature=None if curv is None else curv[:m].copy(), voxel_of_point=vox[: c.n])


def box_dedup(src, tgt, radius: float, ctx: Context | None = None) -> np.ndarray:
    """accumulate_clouds.cpp:100-111: indices of the source points that lie in no target point's
    +/- radius box (the points the tool keeps before its SOR pass)."""
    ctx = ctx or default_context()
    s, t = _hc(src), _hc(tgt)
    kept = np.empty(max(s.n, 1), dtype=np.int32)
    cnt = C.c_int64(0)
    ctx._check(ctx._lib.lc3d_box_dedup(ctx._h, s.ref(), t.ref(), float(radius), kept.ctypes.data,
                                       C.byref(cnt)), "lc3d_box_dedup")
    return kept[: cnt.value].copy()


def euclidean_clusters(cloud, tolerance: float, min_size: int, max_size: int, ctx: Context | None = None):
    """cluster_extraction.cpp:88-101 (pcl::EuclideanClusterExtraction::extract): returns
    (labels, sizes) — labels[i] = rank of point i's cluster (size descending) or -1."""
    ctx = ctx or default_context()
    c = _hc(cloud)
    labels = np.empty(max(c.n, 1), dtype=np.int32)
    sizes = np.zeros(max(c.n, 1), dtype=np.int64)
    cnt = C.c_int64(0)
    ctx._check(ctx._lib.lc3d_euclidean_clusters(ctx._h, c.ref(), float(tolerance), int(min_size), int(max_size),
                                                labels.ctypes.data, sizes.ctypes.data, sizes.size,
                                                C.byref(cnt)), "lc3d_euclidean_clusters")
    return labels[: c.n].copy(), sizes[: cnt.value].copy()


def transform(cloud, T, ctx: Context | None = None):
    ctx = ctx or default_context()
    c = _hc(cloud)
    Tm = (C.c_float * 16)(*np.asarray(T, dtype=np.float32).reshape(16))
    xyz = np.empty((c.n, 3), dtype=np.float32)
    nrm = np.empty((c.n, 3), dtype=np.float32) if c.normal is not None else None
    ctx._check(ctx._lib.lc3d_transform(ctx._h, c.ref(), Tm, xyz.ctypes.data,
                                       None if nrm is None else nrm.ctypes.data), "lc3d_transform")
    return xyz, nrm


# ---------------------------------------------------------------- PCL-named classes ----

class IterativeClosestPoint:
    """pcl::IterativeClosestPoint<PointT,PointT> as used at fine_registration.cpp:105-126."""
    _mode = POINT_TO_POINT

    def __init__(self, ctx: Context | None = None):
        self._ctx = ctx
        self._src = self._tgt = None
        self._max_corr = float(np.sqrt(np.finfo(np.float64).max))  # PCL default: sqrt(DBL_MAX)
        self._max_iter = 10
        self._teps = 0.0
        self._feps = -np.finfo(np.float64).max
        self._res = None

    def setInputSource(self, cloud):
        self._src = _hc(cloud)

    def setInputTarget(self, cloud):
        self._tgt = _hc(cloud)

    def setMaxCorrespondenceDistance(self, d):
        self._max_corr = float(d)

    def setMaximumIterations(self, n):
        self._max_iter = int(n)

    def setTransformationEpsilon(self, e):
        self._teps = float(e)

    def setEuclideanFitnessEpsilon(self, e):
        self._feps = float(e)

    def align(self):
        """Returns the registered cloud (xyz[, normals]) like icp.align(*registered)."""
        if self._src is None or self._tgt is None:
            raise Lc3dError("No input source/target given")
        self._res = icp_align(self._src, self._tgt, self._max_corr, self._max_iter, self._teps, self._feps,
                              mode=self._mode, compute_fitness=True, want_registered=True, ctx=self._ctx)
        return self._res["registered_xyz"], self._res.get("registered_normal")

    def getFinalTransformation(self):
        return self._res["transformation"]

    def hasConverged(self):
        return self._res["converged"]

    def getFitnessScore(self):
        return self._res["fitness"]


class IterativeClosestPointWithNormals(IterativeClosestPoint):
    """pcl::IterativeClosestPointWithNormals (TransformationEstimationPointToPlaneLLS)."""
    _mode = POINT_TO_PLANE


class NormalEstimation:
    """pcl::NormalEstimation as used at normal_estimation.cpp:84-108."""

    def __init__(self, ctx: Context | None = None):
        self._ctx, self._cloud, self._k, self._vp = ctx, None, 0, (0.0, 0.0, 0.0)

    def setInputCloud(self, cloud):
        self._cloud = _hc(cloud)

    def setKSearch(self, k):
        self._k = int(k)

    def setViewPoint(self, x, y, z):
        self._vp = (float(x), float(y), float(z))

    def useSensorOriginAsViewPoint(self):
        self._vp = (0.0, 0.0, 0.0)  # plain PLY clouds carry a zero sensor origin

    def compute(self):
        return normals(self._cloud, self._k, self._vp, ctx=self._ctx)


class VoxelGrid:
    """pcl::VoxelGrid as used at cloud_downsampling.cpp:73-76."""

    def __init__(self, ctx: Context | None = None):
        self._ctx, self._cloud, self._leaf = ctx, None, (1.0, 1.0, 1.0)

    def setInputCloud(self, cloud):
        self._cloud = _hc(cloud)

    def setLeafSize(self, lx, ly, lz):
        self._leaf = (float(lx), float(ly), float(lz))

    def filter(self):
        return voxel_grid(self._cloud, self._leaf, ctx=self._ctx)


class EuclideanClusterExtraction:
    """pcl::EuclideanClusterExtraction as used at cluster_extraction.cpp:94-101."""

    def __init__(self, ctx: Context | None = None):
        self._ctx, self._cloud, self._tol, self._min, self._max = ctx, None, 0.0, 1, 2 ** 31 - 1

    def setInputCloud(self, cloud):
        self._cloud = _hc(cloud)

    def setClusterTolerance(self, t):
        self._tol = float(t)

    def setMinClusterSize(self, n):
        self._min = int(n)

    def setMaxClusterSize(self, n):
        self._max = int(n)

    def setSearchMethod(self, tree=None):  # the grid index replaces the kd-tree
        pass

    def extract(self):
        """List of index arrays (ascending indices), largest cluster first."""
        labels, sizes = euclidean_clusters(self._cloud, self._tol, self._min, self._max, ctx=self._ctx)
        order = np.argsort(labels, kind="stable")
        order = order[labels[order] >= 0]
        return np.split(order.astype(np.int32), np.cumsum(sizes)[:-1]) if len(sizes) else []


class StatisticalOutlierRemoval:
    """pcl::StatisticalOutlierRemoval as used at outlier_removal.cpp:80-93."""

    def __init__(self, ctx: Context | None = None):
        self._ctx, self._cloud, self._k, self._mul, self._neg = ctx, None, 1, 0.0, False

    def setInputCloud(self, cloud):
        self._cloud = _hc(cloud)

    def setMeanK(self, k):
        self._k = int(k)

    def setStddevMulThresh(self, m):
        self._mul = float(m)

    def setNegative(self, neg):
        self._neg = bool(neg)

    def filter(self):
        """Returns the kept indices (input order)."""
        return sor(self._cloud, self._k, self._mul, self._neg, ctx=self._ctx)[0]
