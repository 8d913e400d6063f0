import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def ctx():
    from scde_b200 import _lib

    if _lib.lib().scde_b200_device_count() < 1:
        pytest.fail("GPU test selected but no CUDA device is visible (no CPU fallback exists)")
    return _lib.default_context(0)
