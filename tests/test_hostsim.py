"""Kernel logic on the CPU: the device headers compiled by g++ with a one-lane warp policy
(tests/hostsim) against the reference fixtures.  This exercises exactly the code the GPU runs
(psi_math.cuh / psi_quad.cuh) plus the product's descriptor builders; the `-m gpu` tests repeat
the comparison through the real library."""
import numpy as np
import pytest

import hostsim_util as hs
from common import carr, det_extended, load, lu_noise_floor, product_desc, resonant_floor
from psitools_public_b200 import TanhSinh


@pytest.fixture(scope="module")
def ts():
    hs.build()
    return TanhSinh()


def test_background(ts):
    for name, g in load("background.json").items():
        b = hs.background(product_desc(g["spec"]), ts)
        assert abs(b["vgx"] - g["vgx"]) <= 2e-14*abs(g["vgx"]), name
        assert abs(b["vgy"] - g["vgy"]) <= 2e-14*abs(g["vgy"]), name


def test_disp_eval(ts):
    bg = load("background.json")
    for case in load("disp.json")["cases"]:
        spec = bg[case["dist"]]["spec"]
        w = carr(case["w"])
        D, M, nodes = hs.disp_eval(product_desc(spec), ts, w, case["kx"], case["kz"], case["alpha"], True)
        ref = carr(case["D"])
        Mref = np.array([carr(m) for m in case["M"]])
        scale = np.median(np.abs(ref))
        c, vgx = spec.get("c", 20.0), bg[case["dist"]]["vgx"]
        for i in range(len(w)):
            has_res = len(case["poles"][i]) > len(bg[case["dist"]]["poles"])
            rt = max(1e-10, resonant_floor(w[i]) if has_res else 0.0)
            # (a) the seven integrals
            assert np.max(np.abs(M[i] - Mref[i])) <= 10*rt*np.max(np.abs(Mref[i])), (case["dist"], w[i])
            # (b) D against the exact determinant of the reference's own entries
            ex = det_extended(Mref[i], w[i], case["kx"], case["kz"], case["alpha"], c, vgx)
            assert abs(D[i] - ex) <= rt*abs(ex) + rt*scale, (case["dist"], w[i])
            # (c) D against the reference's LU determinant, within its rounding noise
            tol = rt*abs(ref[i]) + rt*scale + lu_noise_floor(w[i], case["kx"], case["kz"], c, vgx)
            assert abs(D[i] - ref[i]) <= tol, (case["dist"], w[i])
        if not spec.get("single", False):
            assert nodes >= 300*len(w)
