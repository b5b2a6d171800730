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


@pytest.fixture()
def hostsim_backend(monkeypatch):
    """Route the product's host classes to the lane-emulation harness instead of the GPU library so
    that their control logic (lock-step engine, PSIMode, schedulers, grid refinement) can be tested
    in the GPU-less container.  Test-only: the product has no such switch."""
    import hostsim_util as hs
    from psitools_public_b200 import _device
    cache = {}

    def fake_get_context(integrator, device=None):
        key = (int(integrator.p), int(integrator.max_m))
        if key not in cache:
            cache[key] = hs.HostsimContext(integrator)
        return cache[key]

    hs.build()
    monkeypatch.setattr(_device, "get_context", fake_get_context)
    return cache


@pytest.fixture(params=["hostsim", pytest.param("gpu", marks=pytest.mark.gpu)])
def backend(request):
    """Host-logic tests run twice: on the CPU lane-emulation backend (always) and on the real CUDA
    library (``-m gpu``)."""
    if request.param == "hostsim":
        return request.getfixturevalue("hostsim_backend")
    request.getfixturevalue("built")
    return "gpu"
