import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # The built artefacts are git-ignored: on a fresh checkout build them once (nvcc cross-compiles
    # without a GPU).  No-op when liblc3d.so and the tools are already there.
    lib = os.path.join(ROOT, "lowcost3dreconstruction_b200", "csrc", "liblc3d.so")
    tool = os.path.join(ROOT, "lowcost3dreconstruction_b200", "tools", "bin", "cluster_extraction")
    if not (os.path.exists(lib) and os.path.exists(tool)):
        import __graft_entry__
        __graft_entry__.build()


@pytest.fixture(scope="session")
def ctx():
    from lowcost3dreconstruction_b200 import api
    c = api.Context(0)
    yield c
    c.close()
