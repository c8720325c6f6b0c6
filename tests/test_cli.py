"""CLI contract of the four drop-in tools.  The argument handling and
failure paths run without a GPU; the end-to-end runs are marked gpu."""
import os
import subprocess

import numpy as np
import pytest

from lowcost3dreconstruction_b200 import synth
from tests import plyutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "lowcost3dreconstruction_b200", "tools", "bin")
TOOLS = ["fine_registration", "normal_estimation", "cloud_downsampling", "outlier_removal", "accumulate_clouds"]


def run(tool, *args):
    p = subprocess.run([os.path.join(BIN, tool), *args], capture_output=True, text=True, timeout=300)
    return p.returncode, p.stdout, p.stderr


@pytest.mark.parametrize("tool", TOOLS)
def test_help_exits_zero(tool):
    rc, out, err = run(tool, "--help")
    assert rc == 0 and "Options:" in out and "-i [ --input ] arg" in out and "-h [ --help ]" in out
    rc2, out2, _ = run(tool, "-h")
    assert rc2 == 0 and out2 == out


def test_help_lists_reference_options_and_defaults():
    _, out, _ = run("fine_registration", "--help")
    for o in ("--distance_threshold arg (=0.10000000000000001)", "--max_iterations arg (=50)",
              "--transformation_epsilon arg (=1.0000000000000001e-09)", "--euclidean_fitness_epsilon arg (=0.001)",
              "-t [ --target ] arg", "-a [ --accumulated ] arg"):
        assert o in out
    _, out, _ = run("normal_estimation", "--help")
    assert out.startswith("Estimate a set of normals for all the points in the input dataset.")
    for o in ("-n [ --neighbors ] arg (=50)", "-r [ --reverse_normals ]", "-c [ --centroid ]", "-z [ --origin ]"):
        assert o in out
    _, out, _ = run("cloud_downsampling", "--help")
    assert "-s [ --leaf_size ] arg (=1)" in out
    _, out, _ = run("outlier_removal", "--help")
    assert "-f [ --outliers_file ]" in out and "-d [ --dev_mult ] arg (=1)" in out


@pytest.mark.parametrize("tool", TOOLS)
def test_missing_mandatory_options(tool):
    rc, out, err = run(tool)
    assert rc == 255  # return -1
    assert err.startswith("Correct mode of use: ") and "-i input.ply" in err and "-o output.ply [opts]" in err
    if tool in ("fine_registration", "accumulate_clouds"):
        assert "-t target.ply" in err


def test_option_errors():
    rc, _, err = run("fine_registration", "--bogus", "1")
    assert rc == 255 and err.strip() == "ERROR: unrecognised option '--bogus'"
    rc, _, err = run("fine_registration", "-i")
    assert rc == 255 and "the required argument for option '--input' is missing" in err
    rc, _, err = run("fine_registration", "-i", "a", "-t", "b", "-o", "c", "--max_iterations", "0")
    assert rc == 255 and err.strip() == "max_iterations needs to be greater than zero."
    rc, _, err = run("normal_estimation", "-i", "a", "-o", "b", "-c", "-z")
    assert rc == 255 and "not possible to use the centroid and origin as viewpoint at the same time" in err
    rc, _, err = run("outlier_removal", "-i", "a", "-o", "b", "--neighbors", "-3")
    assert rc == 255 and "is invalid" in err
    # unambiguous long prefixes are accepted, as with boost's default style
    rc, _, err = run("fine_registration", "--inp", "/nonexistent.ply", "--tar", "x", "--out", "y", "--dist", "0.02")
    assert rc == 255 and err.strip() == "Couldn't load input cloud file"


def test_load_failures(tmp_path):
    rc, _, err = run("cloud_downsampling", "-i", "/nonexistent.ply", "-o", str(tmp_path / "o.ply"))
    assert rc == 255 and err.strip() == "Couldn't load input point cloud: /nonexistent.ply"
    good = str(tmp_path / "g.ply")
    plyutil.write_capture_ascii(good, np.zeros((3, 3)))
    rc, out, err = run("fine_registration", "-i", good, "-t", "/nonexistent.ply", "-o", str(tmp_path / "o.ply"))
    assert rc == 255 and err.strip() == "Couldn't load input target file"
    assert out.startswith("Loaded 3 data points from " + good)


# ------------------------------------------------------------------------------- GPU runs

@pytest.fixture(scope="module")
def ply_pair(tmp_path_factory):
    d = tmp_path_factory.mktemp("ply")
    # the capture tool's ASCII PLY keeps 6 decimals: the tools see exactly these values
    tgt = np.round(synth.kinect_view(0, scale=0.2, backdrop="panel").astype(np.float64), 6).astype(np.float32)
    src = np.round(synth.kinect_view(1, scale=0.2, backdrop="panel").astype(np.float64), 6).astype(np.float32)
    plyutil.write_capture_ascii(str(d / "src.ply"), src)
    plyutil.write_capture_ascii(str(d / "tgt.ply"), tgt)
    return d, src, tgt


@pytest.mark.gpu
def test_pipeline_end_to_end(ply_pair, ctx):
    from lowcost3dreconstruction_b200 import api
    d, src, tgt = ply_pair
    s, t = str(d / "src.ply"), str(d / "tgt.ply")
    # outlier_removal in place (scripts/alignment.sh:99 uses -i F -o F), with the outliers file
    t2 = str(d / "tgt_f.ply")
    rc, out, err = run("outlier_removal", "-i", t, "-o", t2, "--neighbors", "20", "--dev_mult", "2.0", "-f")
    assert rc == 0, err
    assert out.startswith("Cloud before filtering: \nheader: seq: 0 stamp: 0 frame_id: \n\npoints[]: %d\nwidth: %d\nheight: 1\nis_dense: 1\n" % (len(tgt), len(tgt)))
    kept, _, _ = api.sor(tgt, 20, 2.0, ctx=ctx)
    f = plyutil.read_pcl_binary(t2)
    assert len(f) == len(kept) and np.array_equal(f["xyz"], tgt[kept])
    o = plyutil.read_pcl_binary(str(d / "tgt_f_outliers.ply"))
    assert len(o) + len(f) == len(tgt)
    # normals
    t3 = str(d / "tgt_n.ply")
    rc, out, err = run("normal_estimation", "-i", t2, "-o", t3, "-n", "20")
    assert rc == 0, err
    nrm, curv = api.normals(tgt[kept], 20, ctx=ctx)
    n = plyutil.read_pcl_binary(t3)
    assert np.array_equal(n["normal"], nrm) and np.array_equal(n["curvature"], curv)
    assert np.array_equal(n["rgb"], np.full((len(kept), 3), 128, np.uint8))
    rc, _, _ = run("normal_estimation", "-i", t2, "-o", str(d / "tgt_nr.ply"), "-n", "20", "-r")
    assert np.array_equal(plyutil.read_pcl_binary(str(d / "tgt_nr.ply"))["normal"], -nrm)
    # fine registration (reference behaviour: point-to-point) + accumulated + matrix file
    reg, acc, mat = str(d / "reg.ply"), str(d / "acc.ply"), str(d / "T.txt")
    rc, out, err = run("fine_registration", "-i", s, "-t", t3, "-o", reg, "-a", acc, "--distance_threshold", "0.02",
                       "--matrix_file", mat)
    assert rc == 0, err
    lines = out.strip().split("\n")
    assert lines[0] == f"Loaded {len(src)} data points from {s}"
    assert lines[1] == f"Loaded {len(kept)} data points from {t3}"
    assert lines[2] in ("Has converged: True", "Has converged: False") and lines[3].startswith("Score: ")
    g = api.icp_align(src, tgt[kept], 0.02, 50, want_registered=True, ctx=ctx)
    M = np.array([[float(v) for v in ln.split()] for ln in lines[4:8]])
    assert np.allclose(M, g["transformation"], rtol=2e-5, atol=1e-7)       # printed with 6 significant digits
    assert abs(float(lines[3].split()[1]) - g["fitness"]) <= 1e-5 * g["fitness"]
    widths = {len(ln) for ln in lines[4:8]}
    assert len(widths) == 1                                                # Eigen-style aligned columns
    from lowcost3dreconstruction_b200 import chain
    assert np.allclose(chain.read_matrix_file(mat), g["transformation"], atol=1e-7)
    r = plyutil.read_pcl_binary(reg)
    assert np.array_equal(r["xyz"], g["registered_xyz"])
    a = plyutil.read_pcl_binary(acc)
    assert len(a) == len(src) + len(kept) and np.array_equal(a["xyz"][len(src):], tgt[kept])
    # point-to-plane extension
    rc, out, err = run("fine_registration", "-i", s, "-t", t3, "-o", reg, "--distance_threshold", "0.02", "--point_to_plane")
    assert rc == 0, err
    # voxel grid
    v = str(d / "vox.ply")
    rc, out, err = run("cloud_downsampling", "-i", t3, "-o", v, "-s", "0.01")
    assert rc == 0 and "Cloud after filtering: " in out
    from lowcost3dreconstruction_b200._capi import HostCloud
    gv = api.voxel_grid(HostCloud(tgt[kept], normal=nrm, curvature=curv,
                                  rgba=np.full(len(kept), 0xff808080, np.uint32)), 0.01, ctx=ctx)
    pv = plyutil.read_pcl_binary(v)
    assert np.array_equal(pv["xyz"], gv["xyz"]) and np.array_equal(pv["rgb"], np.full((len(pv), 3), 128, np.uint8))


def test_chain_registration_cli_contract():
    rc, out, err = run("chain_registration", "--help")
    assert rc == 0 and "-n [ --num_captures ] arg" in out and "--point_to_plane" in out
    rc, out, err = run("chain_registration")
    assert rc == 255 and err.startswith("Correct mode of use: ")
    rc, out, err = run("chain_registration", "-n", "3", "-d", "/nonexistent")
    assert rc == 255 and err.strip() == "Couldn't load input point cloud: /nonexistent/0.ply"


@pytest.mark.gpu
def test_chain_registration_matches_python_chain(tmp_path, ctx):
    """The in-process chain tool gives the same pair transforms / composed poses as
    chain.register_chain over the C ABI, and writes transform -t readable matrices."""
    from lowcost3dreconstruction_b200 import api, chain
    n = 4
    views = []
    for v in range(n):
        c = np.round(synth.kinect_view(v, step_deg=4.0, scale=0.15, backdrop="panel").astype(np.float64), 6).astype(np.float32)
        views.append(c)
        plyutil.write_capture_ascii(str(tmp_path / f"{v}.ply"), c)
    out = tmp_path / "out"
    out.mkdir()
    rc, so, err = run("chain_registration", "-n", str(n), "-d", str(tmp_path), "-o", str(out), "--distance_threshold", "0.03")
    assert rc == 0, err
    ref = chain.register_chain(n, lambda v: views[v], lambda s, t: api.icp_align(s, t, 0.03, 50, ctx=ctx))
    for i in range(1, n):
        G = chain.read_matrix_file(str(out / f"fine_{i}.txt"))
        assert np.allclose(G, ref["pose"][i], atol=2e-6)
        moved = plyutil.read_pcl_binary(str(out / f"{i}.ply"))["xyz"]
        exp, _ = api.transform(views[i], ref["pose"][i].astype(np.float32), ctx=ctx)
        assert np.abs(moved - exp).max() < 1e-5
    assert np.array_equal(plyutil.read_pcl_binary(str(out / "0.ply"))["xyz"], views[0])
    assert so.count("Has converged: ") == n - 1


def test_accumulate_clouds_all_needs_no_gpu(tmp_path):
    """--all (the only mode scripts/integration.sh:98 uses) is a plain concatenation."""
    a, b, o = str(tmp_path / "a.ply"), str(tmp_path / "b.ply"), str(tmp_path / "o.ply")
    xa, xb = np.arange(9.0).reshape(3, 3), np.arange(12.0).reshape(4, 3) + 100
    plyutil.write_capture_ascii(a, xa)
    plyutil.write_capture_ascii(b, xb)
    rc, out, err = run("accumulate_clouds", "--all", "-i", a, "-t", b, "-o", o)
    assert rc == 0, err
    assert out.startswith(f"Loaded 3 data points from {a}\nLoaded 4 data points from {b}\nCloud before accumulate: ")
    assert "Cloud after accumulate: \nheader: seq: 0 stamp: 0 frame_id: \n\npoints[]: 7\n" in out
    pts = plyutil.read_pcl_binary(o)
    assert np.array_equal(pts["xyz"], np.concatenate([xb, xa]).astype(np.float32))  # target first, then source
    rc, _, err = run("accumulate_clouds", "-i", a, "-t", "/nonexistent.ply", "-o", o)
    assert rc == 255 and err.strip() == "Couldn't load target point cloud: /nonexistent.ply"


@pytest.mark.gpu
def test_accumulate_clouds_dedup(ply_pair, ctx):
    from lowcost3dreconstruction_b200 import api
    d, src, tgt = ply_pair
    o = str(d / "acc_dedup.ply")
    rc, out, err = run("accumulate_clouds", "-i", str(d / "src.ply"), "-t", str(d / "tgt.ply"), "-o", o,
                       "--radius", "0.004", "--clean_neighbors", "10", "--dev_mult", "2.0", "--negative")
    assert rc == 0, err
    kept = api.box_dedup(src, tgt, 0.004, ctx=ctx)
    kept2, _, _ = api.sor(src[kept], 10, 2.0, ctx=ctx)
    want = src[kept][kept2]
    pts = plyutil.read_pcl_binary(o)
    assert len(pts) == len(tgt) + len(want)
    assert np.array_equal(pts["xyz"][: len(tgt)], tgt) and np.array_equal(pts["xyz"][len(tgt):], want)
    assert np.array_equal(plyutil.read_pcl_binary(str(d / "acc_dedup_negative.ply"))["xyz"], want)
    assert "|----|----|" in out  # the progress banner the reference prints on stdout


def test_cluster_extraction_cli_contract(tmp_path):
    rc, out, _ = run("cluster_extraction", "--help")
    assert rc == 0 and out.startswith("Euclidean cluster extraction.")
    for o in ("-f [ --outliers_file ]", "-p [ --cluster_percentage ] arg (=0.25)",
              "-t [ --tolerance ] arg (=0.02)"):
        assert o in out
    rc, _, err = run("cluster_extraction")
    assert rc == 255 and err.startswith("Correct mode of use: ") and err.strip().endswith("-i input.ply -o output.ply")
    rc, _, err = run("cluster_extraction", "-i", "a", "-o", "b", "-p", "1.5")
    assert rc == 255 and err.strip() == "cluster_percentage must be a value between 0 and 1"
    rc, _, err = run("cluster_extraction", "-i", "/nonexistent.ply", "-o", str(tmp_path / "o.ply"))
    assert rc == 255 and err.strip() == "Couldn't load input point cloud: /nonexistent.ply"


@pytest.mark.gpu
def test_cluster_extraction_matches_oracle(tmp_path):
    from oracle import oracle as orc
    rng = np.random.default_rng(11)
    a = (rng.normal(size=(4000, 3)) * 0.03).astype(np.float32)
    b = (rng.normal(size=(2500, 3)) * 0.03).astype(np.float32) + np.array([0.6, 0, 0], np.float32)
    c = rng.uniform(-1, 1, size=(300, 3)).astype(np.float32)
    pts = np.round(np.concatenate([a, b, c]).astype(np.float64), 6)
    pts = pts[rng.permutation(len(pts))]
    i, o = str(tmp_path / "in.ply"), str(tmp_path / "out.ply")
    plyutil.write_capture_ascii(i, pts)
    x = pts.astype(np.float32)
    rc, out, err = run("cluster_extraction", "-i", i, "-o", o, "-t", "0.02", "-p", "0.2", "-f")
    assert rc == 0, err
    labels, sizes = orc.euclidean_clusters(x, 0.02, int(len(x) * 0.2), len(x))
    assert len(sizes) == 2 and f"{len(sizes)} cluster(s) extracted." in out
    want = np.concatenate([np.flatnonzero(labels == r) for r in range(len(sizes))])
    assert np.array_equal(plyutil.read_pcl_binary(o)["xyz"], x[want])
    assert np.array_equal(plyutil.read_pcl_binary(str(tmp_path / "out_outliers.ply"))["xyz"], x[labels < 0])
    # nothing large enough: the reference's error path
    rc, out, err = run("cluster_extraction", "-i", i, "-o", o, "-t", "0.0001", "-p", "0.9")
    assert rc == 255 and err.strip() == "Could not extact clusters for the given dataset"


# ------------------------------------------------------------------ transform

def test_transform_cli_contract(tmp_path):
    """pcl_tools/transform.cpp: options, usage line, matrix-file errors (no GPU needed)."""
    rc, out, _ = run("transform", "--help")
    assert rc == 0 and out.startswith("Transforms Point cloud.\n\nOptions:")
    assert "-t [ --transform ] arg" in out and "File containing a 4x4 transformation matrix" in out
    rc, _, err = run("transform", "-i", "a.ply", "-o", "b.ply")
    assert rc == 255 and err.startswith("Correct mode of use: ")
    assert err.strip().endswith("-i input.ply -o output.ply -t transform_file.txt")
    rc, _, err = run("transform", "-i", "/nonexistent.ply", "-o", "b.ply", "-t", "m.txt")
    assert rc == 255 and err.strip() == "Couldn't load input point cloud: /nonexistent.ply"
    pts = synth.kinect_view(0, scale=0.05)
    src = str(tmp_path / "in.ply")
    plyutil.write_capture_ascii(src, pts)
    rc, _, err = run("transform", "-i", src, "-o", str(tmp_path / "o.ply"), "-t", str(tmp_path / "missing.txt"))
    assert rc == 255 and err.strip() == "Unable to open file: " + str(tmp_path / "missing.txt")
    bad = tmp_path / "bad.txt"
    bad.write_text("1 0 0 0\n0 1 0 0\n0 0 1\n")
    rc, _, err = run("transform", "-i", src, "-o", str(tmp_path / "o.ply"), "-t", str(bad))
    assert rc == 255 and err.strip() == "Error on read transform file: " + str(bad)


@pytest.mark.gpu
def test_transform_and_chain_pipeline_end_to_end(tmp_path, ctx):
    from lowcost3dreconstruction_b200 import api, chain
    # transform: matrix file in, points + normals out, equal to the library call
    pts = synth.kinect_view(0, scale=0.25, backdrop="none")
    nrm, curv = api.normals(pts, 15, ctx=ctx)
    src, dst, mat = str(tmp_path / "in.ply"), str(tmp_path / "out.ply"), str(tmp_path / "T.txt")
    plyutil.write_capture_ascii(src, pts, normals=nrm)
    T = synth.rigid(3.0, -8.0, 1.5, [0.01, -0.02, 0.03])
    chain.write_matrix_file(mat, T)
    rc, _, err = run("transform", "-i", src, "-o", dst, "-t", mat)
    assert rc == 0, err
    loaded = plyutil.read_pcl_binary(src if False else dst)
    inp_xyz = np.round(pts.astype(np.float64), 6).astype(np.float32)  # the ASCII writer keeps 6 decimals
    inp_nrm = np.round(nrm.astype(np.float64), 6).astype(np.float32)
    from lowcost3dreconstruction_b200._capi import HostCloud
    gx, gn = api.transform(HostCloud(inp_xyz, normal=inp_nrm), chain.read_matrix_file(mat).astype(np.float32), ctx=ctx)
    assert np.array_equal(loaded["xyz"], gx) and np.array_equal(loaded["normal"], gn)
    # chain_registration with the in-process pipeline == prepare_view + resident ICP through the API
    views = [synth.apply_transform(synth.turntable_prior(v, 10.0), synth.kinect_view(v, step_deg=10.0, scale=0.5, backdrop="none"))
             for v in range(3)]
    d = tmp_path / "views"
    d.mkdir()
    for v, c in enumerate(views):
        plyutil.write_capture_ascii(str(d / f"{v}.ply"), c)
    o = tmp_path / "reg"
    o.mkdir()
    rc, out, err = run("chain_registration", "-n", "3", "-d", str(d), "-o", str(o), "--distance_threshold", "0.02",
                       "--point_to_plane", "--leaf_size", "0.004", "--neighbors", "20", "--dev_mult", "2.0", "--normals", "15")
    assert rc == 0, err
    loaded = [np.array(plyutil.read_pcl_binary(str(d / f"{v}.ply"))["xyz"]) if False else
              np.round(c.astype(np.float64), 6).astype(np.float32) for v, c in enumerate(views)]
    prep = [api.prepare_view(c, 0.004, 20, 2.0, 15, ctx=ctx)[0] for c in loaded]
    G = np.eye(4)
    first = plyutil.read_pcl_binary(str(o / "0.ply"))
    assert np.array_equal(first["xyz"], prep[0].download()[0])
    for v in (1, 2):
        r = api.icp_align(prep[v], prep[v - 1], 0.02, 50, mode=api.POINT_TO_PLANE, ctx=ctx)
        G = G @ r["transformation"].astype(np.float64)
        assert np.allclose(chain.read_matrix_file(str(o / f"fine_{v}.txt")), G, atol=1e-7)
        x, n, _ = prep[v].download()
        gx, gn = api.transform(HostCloud(x, normal=n), G.astype(np.float32), ctx=ctx)
        w = plyutil.read_pcl_binary(str(o / f"{v}.ply"))
        assert np.array_equal(w["xyz"], gx) and np.array_equal(w["normal"], gn)
