"""Independent physics cross-check: the roots PSIMode finds for the
continuous size distribution against the goldens of the reference's DIRECT eigen-solver.

The direct solver (reference direct.py:664-730, `Converger*`) discretises the size distribution on
N, 2N, 4N collocation sizes and solves a (4N+4)^2 matrix eigenproblem -- no quadrature, no rational
approximation, no contour: nothing in common with the D(omega) path but the physics.  Its tests
(reference test_direct.py) pin the fastest-growing eigenvalue at three successive refinements; the
sequence converges like 1/N towards the continuum limit, which is what PSIMode solves for.  The
goldens below are copied VALUES from those tests (test_direct.py:109-127 PB_0, :205-225 PBT_0,
:281-296 PL_1; the inviscid cases -- the dust-pressure/diffusion cases solve a different model).

Checked here: the PSIMode root lies (a) within 2e-4 of the Richardson limit of the direct sequence
(x3 + (x3-x2), exact for an error ~ 1/N) for the two C0 densities and within 2e-5 of the last
refinement for PowerBumpTail (its kink at aBT spoils the monotone convergence), and (b) closer to it
than the finest direct refinement is to the limit, for the monotone cases."""
import numpy as np
import pytest

DIRECT_GOLDENS = {
    # name: (three refinements of the fastest eigenvalue, Kx, Kz)
    "PL_1": ((0.5512363292041821 + 0.11242857992948893j, 0.5489690582185277 + 0.11216071905699404j,
              0.5478173144269354 + 0.11203395665778133j), 80.0, 100.0),
    "PB_0": ((-0.04298854605814622 + 0.13764347038117714j, -0.04401630203296608 + 0.13752657929813472j,
              -0.04454892638085971 + 0.1374699433910088j), 4.0, 100.0),
    "PBT_0": ((-0.04153265004675379 + 0.13846049287817908j, -0.041966635201637514 + 0.13811256698525176j,
               -0.0418667083429648 + 0.13817651254264796j), 4.0, 100.0),
}


def product_mode(name):
    from psitools_public_b200 import PowerBump, PowerBumpTail, PSIMode, SizeDensity, TanhSinh
    ts = TanhSinh()
    if name == "PL_1":
        return PSIMode(3.0, [1e-8, 1e-1], [-2, 2], [1e-8, 1.0], n_sample=20, tanhsinh_integrator=ts)
    if name == "PB_0":
        pb = PowerBump(amin=1e-8, aP=1.0, aL=2.0/3.0, aR=1.56, bumpfac=2.0)
    else:
        pb = PowerBumpTail(amin=1e-8, aBT=1e-4, aP=1.0, aL=2.0/3.0, aR=1.56, bumpfac=2.0)
    sd = SizeDensity(pb.sigma0, [1e-8, 1.56])
    sd.poles = [pb.get_discontinuity()]
    return PSIMode(3.0, [1e-8, 1.56], [-2, 2], [1e-8, 1.0], n_sample=20, size_density=sd,
                   tanhsinh_integrator=ts)


def oracle_mode(name):
    from oracle import psi_oracle as po
    if name == "PL_1":
        dd = po.DustDist.power_law(3.0, [1e-8, 1e-1])
    else:
        dd = po.DustDist.power_bump(3.0, 1e-8, 1.0, aL=2.0/3.0, aR=1.56, bumpfac=2.0,
                                    aBT=None if name == "PB_0" else 1e-4)
    return po.Mode(po.Dispersion(dd, po.NodeTable()), [-2, 2], [1e-8, 1.0], n_sample=20)


def check_against_direct(name, roots):
    (x1, x2, x3), _, _ = DIRECT_GOLDENS[name]
    assert len(roots) >= 1, name
    best = roots[np.argmax(np.imag(roots))]                    # the fastest-growing mode
    if name == "PBT_0":
        assert abs(best - x3) < 2e-5, (name, best, x3)
        return
    ratio = (x3 - x2)/(x2 - x1)
    assert abs(ratio - 0.5) < 0.03, ratio                      # the direct sequence converges like 1/N
    limit = x3 + (x3 - x2)
    assert abs(best - limit) < 2e-4, (name, best, limit)
    assert abs(best - limit) < 0.2*abs(x3 - limit), (name, best, limit, x3)


@pytest.mark.parametrize("name", sorted(DIRECT_GOLDENS))
def test_oracle_against_direct_solver(name):
    np.random.seed(0)
    _, kx, kz = DIRECT_GOLDENS[name]
    check_against_direct(name, np.asarray(oracle_mode(name).calculate(kx, kz)))


@pytest.mark.parametrize("name", sorted(DIRECT_GOLDENS))
def test_psimode_against_direct_solver(backend, name):
    np.random.seed(0)
    _, kx, kz = DIRECT_GOLDENS[name]
    roots = np.asarray(product_mode(name).calculate(kx, kz))
    check_against_direct(name, roots)
    np.random.seed(0)
    ref = np.asarray(oracle_mode(name).calculate(kx, kz))
    best, rbest = roots[np.argmax(roots.imag)], ref[np.argmax(ref.imag)]
    assert abs(best - rbest) <= 1e-8*abs(rbest), (best, rbest)
