import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def have_gpu() -> bool:
    try:
        from foundation_b200 import pt
        t = pt.PathTracer(8, 8)
        t.close()
        return True
    except Exception:
        return False


@pytest.fixture(scope="session")
def gpu():
    if not have_gpu():
        pytest.fail("no CUDA device / libfoundation_pt.so unusable: GPU tests must not silently pass")
    return True
