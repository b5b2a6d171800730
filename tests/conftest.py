import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "refcheck: imports the reference from /root/reference (skipped when absent)")


@pytest.fixture(scope="session")
def built():
    import __graft_entry__ as ge
    ge.build()
    return True
