"""ctypes binding of include/lc3d.h (the C ABI of liblc3d.so).

The shared library is built in-tree by ``__graft_entry__.build()`` (or
``make -C lowcost3dreconstruction_b200/csrc``).  There is no CPU fallback: if the
library is missing, loading raises; if no CUDA device is usable, ``lc3d_create`` fails.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LC3D_LIB") or os.path.join(_HERE, "csrc", "liblc3d.so")


class Cloud(C.Structure):
    """struct lc3d_cloud (include/lc3d.h)."""
    _fields_ = [
        ("n", C.c_int64),
        ("xyz", C.c_void_p), ("xyz_stride", C.c_int64),
        ("normal", C.c_void_p), ("normal_stride", C.c_int64),
        ("rgba", C.c_void_p), ("rgba_stride", C.c_int64),
        ("curvature", C.c_void_p), ("curvature_stride", C.c_int64),
    ]


class IcpParams(C.Structure):
    """struct lc3d_icp_params."""
    _fields_ = [
        ("max_correspondence_distance", C.c_double),
        ("transformation_epsilon", C.c_double),
        ("euclidean_fitness_epsilon", C.c_double),
        ("max_iterations", C.c_int32),
        ("mode", C.c_int32),
        ("compute_fitness", C.c_int32),
        ("dump_iteration", C.c_int32),
    ]


class IcpResult(C.Structure):
    """struct lc3d_icp_result."""
    _fields_ = [
        ("transformation", C.c_float * 16),
        ("fitness", C.c_double),
        ("last_mse", C.c_double),
        ("last_correspondences", C.c_int64),
        ("converged", C.c_int32),
        ("iterations", C.c_int32),
        ("state", C.c_int32),
        ("reserved", C.c_int32),
        ("ms_upload", C.c_float),
        ("ms_index", C.c_float),
        ("ms_loop", C.c_float),
        ("ms_fitness", C.c_float),
        ("ms_download", C.c_float),
        ("ms_total", C.c_float),
    ]


class IcpOutputs(C.Structure):
    """struct lc3d_icp_outputs."""
    _fields_ = [
        ("registered_xyz", C.c_void_p),
        ("registered_normal", C.c_void_p),
        ("corr_index", C.c_void_p),
        ("corr_dist2", C.c_void_p),
    ]


class PrepareParams(C.Structure):
    """struct lc3d_prepare_params."""
    _fields_ = [
        ("leaf_size", C.c_float),
        ("sor_mean_k", C.c_int32),
        ("sor_stddev_mul", C.c_double),
        ("normals_k", C.c_int32),
        ("viewpoint", C.c_float * 3),
    ]


POINT_TO_POINT = 0
POINT_TO_PLANE = 1
STATE_NAMES = {0: "NOT_CONVERGED", 1: "ITERATIONS", 2: "TRANSFORM", 3: "ABS_MSE", 4: "REL_MSE",
               5: "NO_CORRESPONDENCES"}

# Every symbol include/lc3d.h declares (tests check the library exports all of them).
SYMBOLS = [
    "lc3d_create", "lc3d_destroy", "lc3d_last_error", "lc3d_version", "lc3d_launch_count",
    "lc3d_debug_grid_info", "lc3d_debug_alloc_count",
    "lc3d_chain_create", "lc3d_chain_destroy", "lc3d_chain_last_error", "lc3d_chain_run",
    "lc3d_cloud_upload", "lc3d_cloud_free", "lc3d_dcloud_size",
    "lc3d_icp_align", "lc3d_icp_align_resident",
    "lc3d_knn", "lc3d_nn", "lc3d_normals", "lc3d_centroid",
    "lc3d_voxel_grid", "lc3d_sor", "lc3d_transform", "lc3d_box_dedup", "lc3d_euclidean_clusters",
    "lc3d_prepare_view", "lc3d_cloud_download", "lc3d_host_register", "lc3d_host_unregister",
    "lc3d_shard_export", "lc3d_shard_connect", "lc3d_icp_align_sharded", "lc3d_shard_close",
]


class HostCloud:
    """Owns numpy arrays and the lc3d_cloud struct that points into them."""

    def __init__(self, xyz, normal=None, rgba=None, curvature=None):
        self.xyz = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
        n = self.xyz.shape[0]
        self.normal = None if normal is None else np.ascontiguousarray(normal, dtype=np.float32).reshape(n, 3)
        self.rgba = None if rgba is None else np.ascontiguousarray(rgba, dtype=np.uint32).reshape(n)
        self.curvature = None if curvature is None else np.ascontiguousarray(curvature, dtype=np.float32).reshape(n)
        c = Cloud()
        c.n = n
        c.xyz, c.xyz_stride = self.xyz.ctypes.data, 12
        if self.normal is not None:
            c.normal, c.normal_stride = self.normal.ctypes.data, 12
        if self.rgba is not None:
            c.rgba, c.rgba_stride = self.rgba.ctypes.data, 4
        if self.curvature is not None:
            c.curvature, c.curvature_stride = self.curvature.ctypes.data, 4
        self.struct = c

    @classmethod
    def from_pcl_aos(cls, aos: np.ndarray) -> "HostCloud":
        """aos: (n,12) float32 view of pcl::PointXYZRGBNormal (48-byte AoS)."""
        self = cls.__new__(cls)
        self.aos = np.ascontiguousarray(aos, dtype=np.float32).reshape(-1, 12)
        base = self.aos.ctypes.data
        c = Cloud()
        c.n = self.aos.shape[0]
        c.xyz, c.xyz_stride = base, 48
        c.normal, c.normal_stride = base + 16, 48
        c.rgba, c.rgba_stride = base + 32, 48
        c.curvature, c.curvature_stride = base + 36, 48
        self.struct = c
        self.xyz = self.aos[:, 0:3]
        self.normal = self.aos[:, 4:7]
        self.rgba = self.aos[:, 8].view(np.uint32)
        self.curvature = self.aos[:, 9]
        return self

    @property
    def n(self) -> int:
        return int(self.struct.n)

    def ref(self):
        return C.byref(self.struct)


def _declare(lib):
    vp, i32, i64, f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    cp = C.POINTER(Cloud)
    lib.lc3d_create.argtypes = [C.c_int, vp, C.POINTER(vp)]
    lib.lc3d_create.restype = C.c_int
    lib.lc3d_destroy.argtypes = [vp]
    lib.lc3d_destroy.restype = None
    lib.lc3d_last_error.argtypes = [vp]
    lib.lc3d_last_error.restype = C.c_char_p
    lib.lc3d_version.argtypes = []
    lib.lc3d_version.restype = C.c_char_p
    lib.lc3d_launch_count.argtypes = [vp]
    lib.lc3d_launch_count.restype = i64
    lib.lc3d_debug_grid_info.argtypes = [vp, C.POINTER(C.c_double)]
    lib.lc3d_debug_grid_info.restype = None
    lib.lc3d_debug_alloc_count.argtypes = []
    lib.lc3d_debug_alloc_count.restype = i64
    lib.lc3d_chain_create.argtypes = [C.c_int, C.c_int32, C.c_int32, C.POINTER(vp)]
    lib.lc3d_chain_create.restype = C.c_int
    lib.lc3d_chain_destroy.argtypes = [vp]
    lib.lc3d_chain_destroy.restype = None
    lib.lc3d_chain_last_error.argtypes = [vp]
    lib.lc3d_chain_last_error.restype = C.c_char_p
    lib.lc3d_chain_run.argtypes = [vp, C.POINTER(Cloud), C.c_int32, C.POINTER(PrepareParams), C.POINTER(IcpParams),
                                   C.POINTER(IcpResult), C.POINTER(i64), C.c_int32]
    lib.lc3d_chain_run.restype = C.c_int
    lib.lc3d_cloud_upload.argtypes = [vp, cp, C.POINTER(vp)]
    lib.lc3d_cloud_upload.restype = C.c_int
    lib.lc3d_cloud_free.argtypes = [vp, vp]
    lib.lc3d_cloud_free.restype = None
    lib.lc3d_dcloud_size.argtypes = [vp]
    lib.lc3d_dcloud_size.restype = i64
    lib.lc3d_icp_align.argtypes = [vp, cp, cp, C.POINTER(IcpParams), C.POINTER(IcpResult), C.POINTER(IcpOutputs)]
    lib.lc3d_icp_align.restype = C.c_int
    lib.lc3d_icp_align_resident.argtypes = [vp, vp, vp, C.POINTER(IcpParams), C.POINTER(IcpResult),
                                            C.POINTER(IcpOutputs)]
    lib.lc3d_icp_align_resident.restype = C.c_int
    lib.lc3d_knn.argtypes = [vp, cp, cp, i32, vp, vp]
    lib.lc3d_knn.restype = C.c_int
    lib.lc3d_nn.argtypes = [vp, cp, cp, f64, vp, vp]
    lib.lc3d_nn.restype = C.c_int
    lib.lc3d_normals.argtypes = [vp, cp, i32, C.POINTER(C.c_float), vp, vp]
    lib.lc3d_normals.restype = C.c_int
    lib.lc3d_centroid.argtypes = [vp, cp, C.POINTER(C.c_float)]
    lib.lc3d_centroid.restype = C.c_int
    lib.lc3d_voxel_grid.argtypes = [vp, cp, C.POINTER(C.c_float), vp, vp, vp, vp, vp, C.POINTER(i64)]
    lib.lc3d_voxel_grid.restype = C.c_int
    lib.lc3d_sor.argtypes = [vp, cp, i32, f64, i32, vp, C.POINTER(i64), vp, C.POINTER(f64)]
    lib.lc3d_sor.restype = C.c_int
    lib.lc3d_box_dedup.argtypes = [vp, cp, cp, f64, vp, C.POINTER(i64)]
    lib.lc3d_box_dedup.restype = C.c_int
    lib.lc3d_euclidean_clusters.argtypes = [vp, cp, f64, i64, i64, vp, vp, i64, C.POINTER(i64)]
    lib.lc3d_euclidean_clusters.restype = C.c_int
    lib.lc3d_transform.argtypes = [vp, cp, C.POINTER(C.c_float), vp, vp]
    lib.lc3d_transform.restype = C.c_int
    lib.lc3d_prepare_view.argtypes = [vp, cp, C.POINTER(PrepareParams), C.POINTER(vp), C.POINTER(i64)]
    lib.lc3d_prepare_view.restype = C.c_int
    lib.lc3d_cloud_download.argtypes = [vp, vp, vp, vp, vp]
    lib.lc3d_cloud_download.restype = C.c_int
    lib.lc3d_host_register.argtypes = [vp, C.c_uint64]
    lib.lc3d_host_register.restype = C.c_int
    lib.lc3d_host_unregister.argtypes = [vp]
    lib.lc3d_host_unregister.restype = C.c_int
    lib.lc3d_shard_export.argtypes = [vp, i64, vp]
    lib.lc3d_shard_export.restype = C.c_int
    lib.lc3d_shard_connect.argtypes = [vp, i32, i32, vp]
    lib.lc3d_shard_connect.restype = C.c_int
    lib.lc3d_icp_align_sharded.argtypes = [vp, vp, vp, C.POINTER(IcpParams), C.POINTER(IcpResult),
                                           C.POINTER(IcpOutputs), C.POINTER(C.c_double)]
    lib.lc3d_icp_align_sharded.restype = C.c_int
    lib.lc3d_shard_close.argtypes = [vp]
    lib.lc3d_shard_close.restype = None
    return lib


_lib = None


def load():
    """Load liblc3d.so (raises if it has not been built — no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)")
        _lib = _declare(C.CDLL(LIB_PATH))
    return _lib
