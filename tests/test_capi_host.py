"""CPU tests of the product's host side: the C-ABI library loads and exports every symbol
include/lc3d.h declares, fails loudly without a GPU (no CPU fallback), and the product never
touches the oracle."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "lc3d.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lc3d_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from lowcost3dreconstruction_b200 import _capi
    lib = _capi.load()
    syms = header_symbols()
    assert len(syms) >= 17
    for s in syms:
        assert hasattr(lib, s), f"liblc3d.so does not export {s}"
    assert sorted(_capi.SYMBOLS) == syms, "SYMBOLS list and include/lc3d.h disagree"
    assert b"sm_100a" in lib.lc3d_version()


def test_library_contains_sm100a_code_only():
    from lowcost3dreconstruction_b200 import _capi
    out = subprocess.run(["cuobjdump", "-lelf", _capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert not re.search(r"sm_(?!100a)\d+", out), out


def test_struct_layouts_match_header():
    from lowcost3dreconstruction_b200 import _capi
    assert C.sizeof(_capi.Cloud) == 72
    assert C.sizeof(_capi.IcpParams) == 40
    assert C.sizeof(_capi.IcpResult) == 64 + 8 + 8 + 8 + 16 + 24
    assert C.sizeof(_capi.IcpOutputs) == 32


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from lowcost3dreconstruction_b200 import api
    with pytest.raises(api.Lc3dError) as e:
        api.Context(0)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)
    with pytest.raises(api.Lc3dError):
        api.icp_align(np.zeros((4, 3), np.float32), np.zeros((4, 3), np.float32))


def test_product_never_uses_the_oracle():
    pkg = os.path.join(ROOT, "lowcost3dreconstruction_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".inc", ".cpp", ".hpp", ".h", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in text.lower() or f == "Makefile" and "oracle" not in text, \
                    f"{os.path.join(dirpath, f)} mentions the oracle"


def test_host_cloud_views():
    from lowcost3dreconstruction_b200._capi import HostCloud
    a = np.zeros((5, 12), np.float32)
    a[:, 0:3] = np.arange(15).reshape(5, 3)
    a[:, 4:7] = 1.0
    hc = HostCloud.from_pcl_aos(a)
    assert hc.n == 5 and hc.struct.xyz_stride == 48 and hc.struct.normal - hc.struct.xyz == 16
    assert hc.struct.rgba - hc.struct.xyz == 32 and hc.struct.curvature - hc.struct.xyz == 36
    hc2 = HostCloud(np.arange(6, dtype=np.float64).reshape(2, 3))
    assert hc2.xyz.dtype == np.float32 and hc2.struct.normal is None


def test_documented_tunables_exist_in_source():
    """INTEGRATION.md's table of LC3D_* environment variables must match what the code reads."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    doc = open(os.path.join(root, "INTEGRATION.md")).read()
    table = doc[doc.index("## Tunables"):doc.index("## Semantics worth knowing")]
    documented = set(re.findall(r"`(LC3D_[A-Z_]+)`", table))
    src = ""
    for d, exts in (("lowcost3dreconstruction_b200/csrc", (".cu", ".cuh", ".inc")), ("lowcost3dreconstruction_b200", (".py",)),
                    ("lowcost3dreconstruction_b200/tools", (".cpp", ".hpp"))):
        for f in os.listdir(os.path.join(root, d)):
            if f.endswith(exts):
                src += open(os.path.join(root, d, f)).read()
    read_by_code = set(re.findall(r'getenv\("(LC3D_[A-Z_]+)"\)', src)) | set(re.findall(r'environ\.get\("(LC3D_[A-Z_]+)"', src))
    assert documented == read_by_code, (sorted(documented - read_by_code), sorted(read_by_code - documented))
