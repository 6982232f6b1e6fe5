import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _has_gpu() -> bool:
    try:
        from msolve_b200.backend import load_library
        return load_library().f4la_b200_device_count() > 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def backend():
    """The product backend; a missing library or GPU is a hard failure for gpu tests."""
    from msolve_b200.backend import Backend
    be = Backend(0)
    yield be
    be.close()


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")
