"""PLY reader / PCL-layout writer of the CLI tools, host only."""
import os
import struct
import subprocess

import numpy as np

from tests import plyutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CONVERT = os.path.join(ROOT, "lowcost3dreconstruction_b200", "tools", "bin", "ply_convert")


def convert(src, dst):
    p = subprocess.run([CONVERT, src, dst], capture_output=True, text=True, timeout=60)
    return p.returncode, p.stdout, p.stderr


def test_capture_ascii_to_pcl_binary(tmp_path):
    rng = np.random.default_rng(0)
    xyz = np.round(rng.normal(0, 100, (500, 3)), 6)
    nrm = np.round(rng.normal(0, 1, (500, 3)), 6)
    rgb = rng.integers(0, 256, (500, 3))
    src, dst = str(tmp_path / "a.ply"), str(tmp_path / "b.ply")
    plyutil.write_capture_ascii(src, xyz, nrm, rgb)
    rc, out, err = convert(src, dst)
    assert rc == 0, err
    assert out.startswith("points 500 normals 1 color 1 curvature 0 dense 1")
    pts = plyutil.read_pcl_binary(dst)  # also checks the header line by line
    assert np.array_equal(pts["xyz"], xyz.astype(np.float32))
    assert np.array_equal(pts["normal"], nrm.astype(np.float32))
    assert np.array_equal(pts["rgb"], rgb.astype(np.uint8))
    assert not pts["curvature"].any()
    # camera element: 19 floats + 2 ints + 2 floats after the vertex block; viewport = (N, 1)
    data = open(dst, "rb").read()
    cam = data[-(17 * 4 + 2 * 4 + 2 * 4):]
    vals = struct.unpack("<17f2i2f", cam)
    assert vals[3:12] == (1, 0, 0, 0, 1, 0, 0, 0, 1) and vals[15:17] == (250.0, 0.5) and vals[17:19] == (500, 1)
    # round trip through our own writer/reader
    rc, out, err = convert(dst, str(tmp_path / "c.ply"))
    assert rc == 0 and np.array_equal(plyutil.read_pcl_binary(str(tmp_path / "c.ply")), pts)


def test_meshlab_style_binary_with_alpha_faces_and_doubles(tmp_path):
    """Arbitrary property order / types, an alpha channel, a face element with list properties,
    an extra scalar — the shapes MeshLab writes after scripts/alignment.sh:100."""
    n = 7
    rng = np.random.default_rng(1)
    xyz = rng.normal(0, 1, (n, 3))
    nrm = rng.normal(0, 1, (n, 3)).astype(np.float32)
    rgba = rng.integers(0, 256, (n, 4)).astype(np.uint8)
    hdr = ("ply\nformat binary_little_endian 1.0\ncomment VCGLIB generated\nelement vertex %d\n"
           "property double x\nproperty double y\nproperty double z\n"
           "property float nx\nproperty float ny\nproperty float nz\n"
           "property uchar red\nproperty uchar green\nproperty uchar blue\nproperty uchar alpha\n"
           "property float quality\n"
           "element face 2\nproperty list uchar int vertex_indices\nend_header\n" % n)
    body = b""
    for i in range(n):
        body += struct.pack("<3d3f4Bf", *xyz[i], *nrm[i], *rgba[i], 0.5)
    body += struct.pack("<B3i", 3, 0, 1, 2) + struct.pack("<B4i", 4, 3, 4, 5, 6)
    src, dst = str(tmp_path / "m.ply"), str(tmp_path / "o.ply")
    open(src, "wb").write(hdr.encode() + body)
    rc, out, err = convert(src, dst)
    assert rc == 0, err
    pts = plyutil.read_pcl_binary(dst)
    assert len(pts) == n and np.array_equal(pts["xyz"], xyz.astype(np.float32))
    assert np.array_equal(pts["normal"], nrm) and np.array_equal(pts["rgb"], rgba[:, :3])


def test_ascii_with_faces_crlf_and_nan(tmp_path):
    src, dst = str(tmp_path / "f.ply"), str(tmp_path / "o.ply")
    open(src, "w").write("ply\r\nformat ascii 1.0\r\ncomment hi\r\nelement vertex 3\r\nproperty float x\r\n"
                         "property float y\r\nproperty float z\r\nelement face 1\r\n"
                         "property list uchar int vertex_index\r\nend_header\r\n"
                         "0 0 0\r\n1 nan 0\r\n0 1 0\r\n3 0 1 2\r\n")
    rc, out, err = convert(src, dst)
    assert rc == 0, err
    assert "points 3" in out and "dense 0" in out
    pts = plyutil.read_pcl_binary(dst)
    assert np.isnan(pts["xyz"][1, 1]) and pts["xyz"][2, 1] == 1.0


def test_rejects_garbage(tmp_path):
    bad = str(tmp_path / "bad.ply")
    open(bad, "w").write("not a ply\n")
    rc, out, err = convert(bad, str(tmp_path / "o.ply"))
    assert rc == 255 and "Couldn't load" in err
    trunc = str(tmp_path / "t.ply")
    open(trunc, "w").write("ply\nformat ascii 1.0\nelement vertex 5\nproperty float x\nproperty float y\n"
                           "property float z\nend_header\n0 0 0\n")
    rc, out, err = convert(trunc, str(tmp_path / "o.ply"))
    assert rc == 255
    be = str(tmp_path / "be.ply")
    open(be, "w").write("ply\nformat binary_big_endian 1.0\nelement vertex 0\nend_header\n")
    rc, out, err = convert(be, str(tmp_path / "o.ply"))
    assert rc == 255 and "unsupported" in err


def test_ascii_number_parser_is_exact(tmp_path):
    """The fast ASCII parser must give the same float32 as a correctly rounded decimal->double->
    float conversion (what PCL's iostream-based reader produces) for every token shape."""
    rng = np.random.default_rng(3)
    toks = ["0", "-0", "+3.", ".5", "-.25", "1e-5", "1E+2", "2.5e3", "123456789.123456789", "0.000001",
            "-0.0000001234567890123456789", "3.4028234e38", "1.17549435e-38", "1e-45", "123456789012345678901234567890",
            "0.1", "0.2", "0.30000000000000004", "9007199254740993", "1.7976931348623157e308", "4.9e-324",
            "inf", "-inf", "nan", "1e400", "-1e-400"]
    for _ in range(3000):
        kind = rng.integers(0, 4)
        x = rng.normal() * 10.0 ** rng.integers(-8, 9)
        toks.append({0: "%.6f" % x, 1: "%.9g" % x, 2: "%.17g" % x, 3: "%e" % x}[int(kind)])
    while len(toks) % 3:
        toks.append("1")
    vals = np.array(toks).reshape(-1, 3)
    src, dst = str(tmp_path / "n.ply"), str(tmp_path / "o.ply")
    with open(src, "w") as f:
        f.write("ply\nformat ascii 1.0\nelement vertex %d\nproperty float x\nproperty float y\nproperty float z\n"
                "end_header\n" % len(vals))
        for r in vals:
            f.write(" ".join(r) + "\n")
    rc, out, err = convert(src, dst)
    assert rc == 0, err
    got = plyutil.read_pcl_binary(dst)["xyz"].reshape(-1)
    with np.errstate(over="ignore"):
        want = np.array([float(t) for t in vals.reshape(-1)], dtype=np.float64).astype(np.float32)
    assert np.array_equal(got.view(np.uint32)[~np.isnan(want)], want.view(np.uint32)[~np.isnan(want)])
    assert np.array_equal(np.isnan(got), np.isnan(want))


def test_parallel_ascii_parse_matches_sequential(tmp_path):
    """Vertex elements of >= 100k points are parsed by several threads (token-count prefix sums
    give every thread a vertex-aligned range): the result must be byte-identical to the
    single-thread parse, whatever the whitespace layout, and the elements after it must still parse."""
    rng = np.random.default_rng(5)
    n = 150_001
    xyz = np.round(rng.normal(0, 300, (n, 3)), 6)
    nrm = np.round(rng.normal(0, 1, (n, 3)), 6)
    rgb = rng.integers(0, 256, (n, 3))
    src = str(tmp_path / "big.ply")
    seps = [" ", "  ", "\t", " \t "]
    with open(src, "w") as f:
        f.write("ply\nformat ascii 1.0\nelement vertex %d\nproperty float x\nproperty float y\nproperty float z\n"
                "property float nx\nproperty float ny\nproperty float nz\nproperty uchar red\nproperty uchar green\n"
                "property uchar blue\nelement face 2\nproperty list uchar int vertex_index\nend_header\n" % n)
        lines = []
        for i in range(n):
            s = seps[i % 4]
            vals = ["%.6f" % v for v in xyz[i]] + ["%.6f" % v for v in nrm[i]] + ["%d" % v for v in rgb[i]]
            eol = "\r\n" if i % 7 == 0 else ("\n\n" if i % 1001 == 0 else "\n")
            # every 5000th vertex is split over two lines: the parser is token-based, not line-based
            if i % 5000 == 17:
                lines.append(s.join(vals[:4]) + "\n" + s.join(vals[4:]) + eol)
            else:
                lines.append(s.join(vals) + eol)
        f.write("".join(lines))
        f.write("3 0 1 2\n3 2 1 0\n")
    outs = {}
    for threads in ("1", "2", "7", "16"):
        dst = str(tmp_path / f"o{threads}.ply")
        p = subprocess.run([CONVERT, src, dst], capture_output=True, text=True, timeout=120,
                           env=dict(os.environ, LC3D_PLY_THREADS=threads))
        assert p.returncode == 0, p.stderr
        assert p.stdout.startswith("points %d normals 1 color 1" % n)
        outs[threads] = open(dst, "rb").read()
    assert outs["1"] == outs["2"] == outs["7"] == outs["16"]
    pts = plyutil.read_pcl_binary(str(tmp_path / "o7.ply"))
    assert np.array_equal(pts["xyz"], xyz.astype(np.float32)) and np.array_equal(pts["rgb"], rgb.astype(np.uint8))
    # truncated body: every thread count reports it
    data = open(src).read()
    open(src, "w").write(data[: len(data) // 2])
    for threads in ("1", "7"):
        p = subprocess.run([CONVERT, src, str(tmp_path / "t.ply")], capture_output=True, text=True, timeout=120,
                           env=dict(os.environ, LC3D_PLY_THREADS=threads))
        assert p.returncode != 0 and "truncated PLY data" in p.stderr
