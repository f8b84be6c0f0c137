import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _native_artifacts():
    """Build the checker (oracle) and, if absent, the product library (nvcc cross-compiles on CPU)."""
    import oracle_binding

    oracle_binding.build_oracle()
    from mcp_raytracer_b200 import _native

    if not os.path.exists(_native.LIB_PATH):
        _native.build()
    yield


def has_gpu() -> bool:
    from mcp_raytracer_b200 import _native

    try:
        return _native.lib().rt_device_count() > 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def gpu():
    if not has_gpu():
        pytest.fail("-m gpu tests need a CUDA device; the render path has no CPU fallback")
    return True
