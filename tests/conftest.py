import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def ctx():
    """A libcrgpu context on cuda:0.  GPU tests FAIL (not skip) when the extension or the GPU is
    missing: there is no CPU fallback to hide behind."""
    from crispresso_b200 import Context
    c = Context(0)
    yield c
    c.close()
