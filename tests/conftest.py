import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests", "hostsim"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def cuda_available():
    try:
        import ctypes
        cudart = ctypes.CDLL("libcudart.so")
    except OSError:
        cudart = None
    import hvb200
    L = hvb200._abi.lib()
    import ctypes as C
    ctx = C.c_void_p()
    xs = np.random.default_rng(0).random((8, 2))
    rc = L.hvb_create(C.byref(ctx), 2, 8, xs.ctypes.data_as(C.c_void_p), 0, None, None, None)
    if rc == 0:
        L.hvb_destroy(ctx)
        return True
    return False


@pytest.fixture(scope="session")
def oracle():
    import hv_oracle
    hv_oracle.build()
    return hv_oracle


@pytest.fixture(scope="session")
def hvb():
    import hvb200
    return hvb200
