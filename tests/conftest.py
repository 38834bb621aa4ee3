import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu():
    try:
        import ctypes
        cuda = ctypes.CDLL("libcuda.so.1")
        n = ctypes.c_int(0)
        return cuda.cuInit(0) == 0 and cuda.cuDeviceGetCount(ctypes.byref(n)) == 0 and n.value > 0
    except OSError:
        return False


HAS_GPU = _has_gpu()


def pytest_collection_modifyitems(config, items):
    if HAS_GPU:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def ps():
    """The ctypes binding with libps_b200.so built (GPU tests must run the native library)."""
    import __graft_entry__ as g
    g.build()
    from ps_b200 import binding
    binding.lib()
    return binding


@pytest.fixture()
def ctx(ps):
    c = ps.Context(0, seed=20261017)
    c.set_fc_precision(ps.PS_FC_FP32)     # the tight tolerances of the parity tests are stated for the exact (FFMA) FcLayer mode; tests of the
    yield c                               # tensor-core modes (the library default is 3xTF32) select them explicitly
    c.close()
