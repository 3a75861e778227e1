import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def built():
    """Build the CUDA library (nvcc cross-compiles without a GPU) and the oracle once."""
    from popscle_b200 import _build
    _build.build_cuda()
    import oracle_py
    oracle_py.build()
    return True


@pytest.fixture(scope="session")
def ctx(built):
    from popscle_b200 import Context
    c = Context(0)  # raises without an sm_100 device: GPU tests must not pass on a fallback
    yield c
    c.close()
