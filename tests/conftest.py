import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def pz():
    import __graft_entry__ as g
    so = os.path.join(g.PKG_DIR, "libpzgpu.so")
    if not os.path.exists(so):
        g.build()
    return g.load_package()
