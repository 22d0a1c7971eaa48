import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


def _has_gpu():
    try:
        import ctypes
        import bazinga_b200 as bz
        lib = bz.load_library()
        n = ctypes.c_int(0)
        return lib.bz_device_count(ctypes.byref(n)) == 0 and n.value > 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def gpu_ctx():
    import bazinga_b200 as bz
    if not _has_gpu():
        pytest.fail("no CUDA device: GPU tests must run on a B200 (there is no CPU fallback)")
    return bz.default_context()
