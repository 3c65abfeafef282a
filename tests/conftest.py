import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Everything is built in-tree once per session (idempotent make)."""
    import __graft_entry__ as g
    need = [os.path.join(ROOT, "longqc_b200", "liblqcov.so"), os.path.join(ROOT, "oracle", "liblqoracle.so"),
            os.path.join(ROOT, "longqc_b200", "csrc", "liblqcov_hostcheck.so")]
    if not all(os.path.exists(p) for p in need):
        g.build()
    yield
