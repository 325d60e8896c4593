import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_sessionstart(session):
    """build the in-tree libraries once if a fresh checkout has none (`make` = __graft_entry__.build(); nvcc cross-compiles
    sm_100a without a GPU).  Existing libraries are left alone: rebuilding after a source change is the caller's `make`."""
    import subprocess
    need = ["slepc_b200/lib/libb200krylov.so", "slepc_b200/lib/libb2kslepc.so", "oracle/_build/liboraclecpu.so"]
    if not all(os.path.exists(os.path.join(ROOT, p)) for p in need):
        r = subprocess.run(["make", "-C", ROOT, "all"], capture_output=True, text=True)
        if r.returncode != 0:
            raise pytest.UsageError("building the libraries failed:\n" + r.stdout[-2000:] + r.stderr[-2000:])


@pytest.fixture(scope="session")
def ctx():
    """b2k context on cuda:0.  No fallback: a missing GPU or library is an error for gpu tests."""
    from slepc_b200 import _b2k
    c = _b2k.Context(0)
    yield c
    c.close()
