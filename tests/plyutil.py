"""Minimal PLY helpers for the CLI tests (independent of the tools' own reader/writer)."""
import numpy as np


def write_capture_ascii(path, xyz, normals=None, rgb=None):
    """The layout capture/depth_capture/depth_capture.cpp:281-308 writes."""
    n = len(xyz)
    normals = np.zeros((n, 3)) if normals is None else normals
    rgb = np.full((n, 3), 128, dtype=int) if rgb is None else rgb
    with open(path, "w") as f:
        f.write("ply\nformat ascii 1.0\nelement vertex %d\n" % n)
        for p in ("x", "y", "z", "nx", "ny", "nz"):
            f.write("property float %s\n" % p)
        for p in ("red", "green", "blue"):
            f.write("property uchar %s\n" % p)
        f.write("end_header\n")
        for i in range(n):
            f.write("%.6f %.6f %.6f %.6f %.6f %.6f %d %d %d\n" % (*xyz[i], *normals[i], *rgb[i]))


def read_pcl_binary(path):
    """Parses the PCL binary layout and checks the header verbatim."""
    data = open(path, "rb").read()
    end = data.index(b"end_header\n") + len(b"end_header\n")
    header = data[:end].decode().split("\n")
    assert header[0] == "ply" and header[1] == "format binary_little_endian 1.0"
    assert header[2] == "comment PCL generated"
    n = int(header[3].split()[-1])
    props = [h.split()[-1] for h in header[4:14]]
    assert props == ["x", "y", "z", "red", "green", "blue", "nx", "ny", "nz", "curvature"], props
    assert header[14] == "element camera 1"
    dt = np.dtype([("xyz", "<f4", 3), ("rgb", "u1", 3), ("normal", "<f4", 3), ("curvature", "<f4")])
    assert dt.itemsize == 31
    pts = np.frombuffer(data, dtype=dt, count=n, offset=end)
    cam = data[end + 31 * n:]
    assert len(cam) == 19 * 4 + 2 * 4 + 2 * 4 - 2 * 4 + 2 * 4 or len(cam) == 17 * 4 + 8 + 8
    return pts
