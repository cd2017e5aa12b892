import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def _have_gpu():
    try:
        import ctypes

        from dcgrid_b200 import _lib
        from dcgrid_b200.params import scene_params

        L = _lib.load()
        h = ctypes.c_void_p()
        p = scene_params(8)
        rc = L.dcg_create_uniform(ctypes.byref(p), 0, ctypes.byref(h))
        if rc == 0:
            L.dcg_destroy(h)
        return rc == 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def gpu():
    if not _have_gpu():
        pytest.fail("this test is marked gpu but no CUDA device / extension is usable (there is no CPU fallback)")
    return True
