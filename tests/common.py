"""Shared helpers for the test-suite: golden fixtures, descriptor builders, tolerances."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
EPS = np.finfo(float).eps


def cplx(p):
    return complex(p[0], p[1])


def carr(ps):
    return np.array([cplx(p) for p in ps], dtype=complex)


def load(name):
    with open(os.path.join(GOLDEN, name)) as fh:
        return json.load(fh)


def DISTS_BY_NAME(name):
    """Distribution spec of a fixture name (the dict make_golden.py's DISTS holds, via background.json)."""
    return load("background.json")[name]["spec"]


def oracle_dist(spec):
    from oracle import psi_oracle as po
    if spec["kind"] == "powerlaw":
        return po.DustDist.power_law(spec["mu"], spec["stokes_range"], spec.get("c", 20.0),
                                     spec.get("power", 3.5), spec.get("single", False))
    return po.DustDist.power_bump(spec["mu"], spec["amin"], spec["aP"], spec["aL"], spec["aR"],
                                  spec["bumpfac"], epstot=spec.get("epstot"), aBT=spec.get("aBT"),
                                  c=spec.get("c", 20.0))


def product_kwargs(spec):
    """PSIDispersion kwargs built from the PRODUCT's own API objects (as a user would)."""
    from psitools_public_b200 import PowerBump, PowerBumpTail, SizeDensity
    kw = dict(dust_to_gas_ratio=spec["mu"], sound_speed_over_eta=spec.get("c", 20.0))
    if spec["kind"] == "powerlaw":
        kw.update(stokes_range=spec["stokes_range"], size_distribution_power=spec.get("power", 3.5),
                  single_size_flag=spec.get("single", False))
        return kw
    if spec["kind"] == "powerbump":
        pb = PowerBump(amin=spec["amin"], aP=spec["aP"], aL=spec["aL"], aR=spec["aR"],
                       bumpfac=spec["bumpfac"], epstot=spec.get("epstot"))
    else:
        pb = PowerBumpTail(amin=spec["amin"], aBT=spec["aBT"], aP=spec["aP"], aL=spec["aL"],
                           aR=spec["aR"], bumpfac=spec["bumpfac"], epstot=spec.get("epstot"))
    sd = SizeDensity(pb.sigma0, [spec["amin"], pb.aR])
    sd.poles = [pb.get_discontinuity()]
    kw.update(stokes_range=[spec["amin"], pb.aR], size_density=sd)
    return kw


def product_desc(spec):
    from psitools_public_b200 import _device
    kw = product_kwargs(spec)
    desc, table = _device.describe_size_density(kw["dust_to_gas_ratio"], kw["stokes_range"],
                                                kw["sound_speed_over_eta"],
                                                kw.get("size_distribution_power", 3.5),
                                                kw.get("single_size_flag", False), kw.get("size_density"))
    assert table is None          # the fixture distributions are analytic families
    return desc


def det_extended(M7, w, kx, kz, alpha, c, vgx):
    """det(P + M) in extended precision with the huge (kx kz c^2/r)^2 products cancelled
    analytically -- the exact determinant of the reference's own matrix entries
    (psi_dispersion.py:189-210, :451), free of LAPACK's LU rounding."""
    LD = np.clongdouble
    w = LD(w)
    kx, kz, c, vgx = (np.longdouble(v) for v in (kx, kz, c, vgx))
    nu = np.longdouble(alpha)*c**2
    r = w - kx*vgx
    d = kx*vgx - w - 1j*nu*(kx**2 + kz**2)
    g = c**2/r
    m = [LD(x) for x in M7]
    e00 = m[0] - 1j*kx**2*nu/3
    e02 = m[2] - 1j*kx*kz*nu/3
    e20 = -1j*kx*kz*nu/3
    e22 = m[6] - 1j*kz**2*nu/3
    minor = (d + e00)*(d + e22) - e02*e20 + g*(kx*kx*(d + e22) + kz*kz*(d + e00) - kx*kz*(e02 + e20))
    a01, a10, a11, a12 = 2j + m[1], -0.5j + m[3], d + m[4], m[5]
    a20 = kx*kz*g + e20
    a22 = kz*kz*g + d + e22
    return complex(a11*minor - a01*(a10*a22 - a12*a20))


def det_terms_scale(M7, w, kx, kz, alpha, c, vgx):
    """Sum of the magnitudes of the terms of det(P + M) in the cancellation-free expansion used by det_extended:
    the natural scale against which an error of D is measured (at a root the terms cancel and |D| itself is no
    yardstick)."""
    nu = alpha*c**2
    r = w - kx*vgx
    d = kx*vgx - w - 1j*nu*(kx**2 + kz**2)
    g = c**2/r
    M7 = np.asarray(M7, dtype=complex)          # (7,) for one point or (n, 7) with w of length n
    m = [M7[..., k] for k in range(7)]
    e00, e02 = m[0] - 1j*kx**2*nu/3, m[2] - 1j*kx*kz*nu/3
    e20, e22 = -1j*kx*kz*nu/3, m[6] - 1j*kz**2*nu/3
    minor = abs((d + e00)*(d + e22)) + abs(e02*e20) + abs(g)*(kx*kx*abs(d + e22) + kz*kz*abs(d + e00)
                                                             + kx*kz*abs(e02 + e20))
    a01, a10, a11, a12 = 2j + m[1], -0.5j + m[3], d + m[4], m[5]
    a20 = kx*kz*g + e20
    a22 = kz*kz*g + d + e22
    out = abs(a11)*minor + abs(a01)*(abs(a10*a22) + abs(a12*a20))
    return float(out) if np.ndim(out) == 0 else out


def lu_noise_floor(w, kx, kz, c, vgx):
    """Magnitude of the rounding noise of an LU determinant of P + M: a few ulps of the products
    (kx kz c^2/r)^2 * d that cancel inside it (see csrc/psi_math.cuh, det_P_plus_M)."""
    r = abs(w - kx*vgx)
    d = abs(kx*vgx - w)
    return 8*EPS*(kx*kz*c*c/r)**2*max(d, 1.0)


def resonant_floor(w):
    """The reference's own reproducibility floor on the resonant near-real-axis band
    (SURVEY.md section 7): relative 3e-17/|Im w|."""
    return 3e-17/max(abs(w.imag), 1e-300)
