"""The C-ABI library builds, loads and exports every symbol include/psi_b200.h declares; without a
GPU it refuses to create a context (no CPU fallback)."""
import ctypes
import os
import re

import pytest

from psitools_public_b200 import _device

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, "include", "psi_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(psi_[a-z0-9_]+)\s*\(", text)))


def test_exports_match_header(built):
    lib = _device.load_library()
    names = header_functions()
    assert names, "no declarations parsed"
    for n in names:
        assert hasattr(lib, n), "missing export " + n
    assert set(names) == set(_device.exported_symbols())


def test_struct_layout_matches_header():
    # psi_dist_desc: 2 ints, 14 doubles, int (+pad), 4 doubles
    assert ctypes.sizeof(_device.DistDesc) == 8 + 14*8 + 8 + 4*8


def test_no_cpu_fallback(built):
    lib = _device.load_library()
    if lib.psi_device_count() > 0:
        pytest.skip("GPU present")
    from psitools_public_b200 import PSIDispersion
    with pytest.raises(_device.PsiError):
        PSIDispersion(3.0, [1e-8, 1e-2])


def test_argument_checks_without_gpu(built):
    """Entry points validate their arguments before touching the device (no compute, no GPU needed)."""
    lib = _device.load_library()
    null = ctypes.c_void_p()
    assert lib.psi_ctx_create(None, 0, None, None, 0, 12, ctypes.c_double(1e-15)) == -2      # PSI_ERR_ARG
    assert b"bad argument" in lib.psi_last_error()
    assert lib.psi_dist_add(null, None) == -2
    assert lib.psi_ctx_sm_count(null) == 0 and lib.psi_ctx_launch_count(null) == 0
    lib.psi_ctx_destroy(null)                                                                # no-op on NULL
    assert lib.psi_disp_eval_host(null, 0, 1, None, None, None, None, None, None, None, None) == -2
    assert lib.psi_contour_batch_host(null, 0, *([None]*13)) == -2
    assert lib.psi_polish_batch_host(null, 0, *([None]*12)) == -2
    assert lib.psi_version().startswith(b"psi_b200")


def test_openblas_discovery():
    a, b = _device.openblas_paths()
    assert os.path.exists(a) and "openblas64_" in a
    assert os.path.exists(b) and "openblas" in b
