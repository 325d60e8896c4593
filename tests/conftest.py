import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def ctx():
    """b2k context on cuda:0.  No fallback: a missing GPU or library is an error for gpu tests."""
    from slepc_b200 import _b2k
    c = _b2k.Context(0)
    yield c
    c.close()
