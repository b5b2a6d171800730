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


@pytest.fixture(params=[1, 32])
def eval_lanes(request):
    """K1's arithmetic on one lane and on 32 emulated lanes (WarpEmu: the lane-strided level loop, the node
    pairs per lane and the butterfly reductions of the warp kernel, run by host threads)."""
    hs.lib().hs_set_eval_lanes(request.param)
    yield request.param
    hs.lib().hs_set_eval_lanes(1)


@pytest.mark.timeout(900)
def test_disp_eval(ts, eval_lanes):
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


def _loewner(Z, F, mask):
    on, off = np.flatnonzero(mask), np.flatnonzero(~mask)
    C = 1.0/(Z[off][:, None] - Z[on][None, :])
    return F[off][:, None]*C - C*F[on][None, :]


@pytest.mark.parametrize("M,m,seed", [(20, 3, 0), (20, 7, 1), (20, 11, 2), (100, 12, 3), (100, 30, 4),
                                      (180, 25, 5), (40, 21, 6)])
def test_aaa_weights_against_numpy_svd(M, m, seed):
    """The device AAA takes its weights from a Householder QR + inverse iteration (csrc/psi_aaa.cuh);
    the reference takes the last right singular vector of numpy's SVD (complex_roots.py:67-82).  On the
    Loewner matrix of a generic function both must give the same vector up to a phase, and the fit
    residual ||A w|| must equal the smallest singular value; for a wide matrix (more support points
    than free samples: M=20, m=11 and M=40, m=21) any null vector is a valid answer."""
    import ctypes as C
    rng = np.random.RandomState(seed)
    Z = rng.uniform(-2, 2, M) + 1j*rng.uniform(1e-3, 1, M)
    poles = rng.uniform(-3, 3, 40) + 1j*rng.uniform(-2, 2, 40)
    res = rng.normal(size=40) + 1j*rng.normal(size=40)
    F = 1e7*(np.sum(res[None, :]/(Z[:, None] - poles[None, :]), axis=1) + np.exp(1j*Z))
    mask = np.zeros(M, dtype=bool)
    mask[rng.choice(M, m, replace=False)] = True
    w = np.zeros(m, dtype=np.complex128)
    resid = C.c_double(0.0)
    lib = hs.lib()
    Zc, Fc, mc = np.ascontiguousarray(Z), np.ascontiguousarray(F), np.ascontiguousarray(mask, dtype=np.int32)
    got_m = lib.hs_aaa_fit(M, Zc.ctypes.data_as(C.c_void_p), Fc.ctypes.data_as(C.c_void_p),
                           mc.ctypes.data_as(C.c_void_p), w.ctypes.data_as(C.c_void_p), C.byref(resid))
    assert got_m == m
    A = _loewner(Z, F, mask)
    sv = np.linalg.svd(A, compute_uv=False)
    assert abs(np.linalg.norm(w) - 1.0) < 1e-12
    fit = np.linalg.norm(A @ w)
    if M - m < m:                      # wide: exact null vector
        assert fit <= 1e-12*sv[0]
        return
    wref = np.linalg.svd(A)[2][-1].conj()
    assert fit <= sv[-1]*(1 + 1e-8) + 1e-15*sv[0]
    gap = (sv[-2] - sv[-1])/sv[0]
    assert abs(abs(np.vdot(wref, w)) - 1.0) < 1e-10 + 1e-14/max(gap, 1e-16), (abs(np.vdot(wref, w)), gap)


@pytest.mark.timeout(300)
@pytest.mark.parametrize("M,m,seed", [(20, 3, 0), (20, 7, 1), (20, 10, 7), (20, 11, 2), (40, 21, 6), (100, 40, 8)])
def test_aaa_fit_on_32_emulated_lanes(M, m, seed):
    """The warp form of the Loewner fit (lane-strided Householder QR, butterfly reductions, and -- for m <= 32 --
    the inverse iteration whose right-hand side lives in registers with the pivot entry broadcast by shuffle)
    run by 32 emulated lanes (tests/hostsim/hostsim.cpp: WarpEmu): (a) the register form and the general
    (memory) form of the triangular solves give BIT-IDENTICAL weights -- the claim behind enabling the
    register form on the device; (b) the weights satisfy the same numpy-SVD checks as the one-lane run;
    (c) for m > 32 the general form is what runs."""
    import ctypes as C
    rng = np.random.RandomState(seed)
    Z = rng.uniform(-2, 2, M) + 1j*rng.uniform(1e-3, 1, M)
    poles = rng.uniform(-3, 3, 40) + 1j*rng.uniform(-2, 2, 40)
    res = rng.normal(size=40) + 1j*rng.normal(size=40)
    F = 1e7*(np.sum(res[None, :]/(Z[:, None] - poles[None, :]), axis=1) + np.exp(1j*Z))
    mask = np.zeros(M, dtype=bool)
    mask[rng.choice(M, m, replace=False)] = True
    lib = hs.lib()
    Zc, Fc, mc = np.ascontiguousarray(Z), np.ascontiguousarray(F), np.ascontiguousarray(mask, dtype=np.int32)
    out = {}
    for general in (0, 1):
        w = np.zeros(m, dtype=np.complex128)
        resid = C.c_double(0.0)
        got_m = lib.hs_aaa_fit_warp(M, Zc.ctypes.data_as(C.c_void_p), Fc.ctypes.data_as(C.c_void_p),
                                    mc.ctypes.data_as(C.c_void_p), w.ctypes.data_as(C.c_void_p),
                                    C.byref(resid), general)
        assert got_m == m
        out[general] = (w, resid.value)
    assert np.array_equal(out[0][0].view(np.float64), out[1][0].view(np.float64))     # bit for bit
    assert out[0][1] == out[1][1]
    w = out[0][0]
    A = _loewner(Z, F, mask)
    sv = np.linalg.svd(A, compute_uv=False)
    assert abs(np.linalg.norm(w) - 1.0) < 1e-12
    fit = np.linalg.norm(A @ w)
    if M - m < m:
        assert fit <= 1e-12*sv[0]
        return
    wref = np.linalg.svd(A)[2][-1].conj()
    assert fit <= sv[-1]*(1 + 1e-8) + 1e-15*sv[0]
    gap = (sv[-2] - sv[-1])/sv[0]
    assert abs(abs(np.vdot(wref, w)) - 1.0) < 1e-10 + 1e-14/max(gap, 1e-16)


@pytest.mark.timeout(600)
def test_mode_search_on_32_emulated_lanes(hostsim_backend):
    """Whole PSIMode searches (psi_mode.cuh: mode_run = sampling, K1, AAA analysis, Ehrlich-Aberth zeros,
    cleanup, zoom domains, secant polish) run by 32 emulated lanes -- the multi-lane code paths of kernels
    K1/K3/K4 (lane-strided loops, butterfly reductions, ballots, shuffle broadcasts) executed on the CPU.
    Against the one-lane run of the same code: same number of roots and of D evaluations, roots equal to
    rounding; against the reference's goldens (test_psi_mode.py:53-67, :136-150): 1e-8."""
    import time
    from psitools_public_b200 import PowerBump, SizeDensity, TanhSinhNoDeepCopy, psi_mode_mpi
    ts = TanhSinhNoDeepCopy()
    pb = PowerBump(amin=1e-8, aP=0.1, aL=2.0/3.0*0.1, aR=0.156, bumpfac=2.0)
    sd = SizeDensity(pb.sigma0, [1e-8, 0.156])
    sd.poles = [pb.get_discontinuity()]
    base = {'real_range': [-2, 2], 'imag_range': [1e-8, 1], 'tanhsinh_integrator': ts}
    items = [
        {'__init__': dict(base, stokes_range=[1e-8, 1e-2], dust_to_gas_ratio=3.0),
         'calculate': {'wave_number_x': 10, 'wave_number_z': 1000}, 'random_seed': 0},
        {'__init__': dict(base, stokes_range=[1e-8, 1.0], dust_to_gas_ratio=3.0),
         'calculate': {'wave_number_x': 0.1, 'wave_number_z': 10}, 'random_seed': 0},
        {'__init__': dict(base, stokes_range=[1e-8, 0.156], dust_to_gas_ratio=10.0, size_density=sd, n_sample=15),
         'calculate': {'wave_number_x': 200, 'wave_number_z': 1000,
                       'guess_roots': [0.3 + 0.05j]}, 'random_seed': 0}]
    out = {}
    try:
        for lanes in (1, 32):
            hs.lib().hs_set_mode_lanes(lanes)
            ms = psi_mode_mpi.MpiScheduler(time.time(), 1e9, verbose=False)
            res = ms.run_all(items)
            assert ms.last_stats["engine"] == "device"
            out[lanes] = ([np.sort_complex(np.asarray(res[i])) for i in range(len(items))], ms.last_stats["evals"])
    finally:
        hs.lib().hs_set_mode_lanes(1)
    assert out[1][1] == out[32][1]                                       # same number of D evaluations
    for a, b in zip(out[1][0], out[32][0]):
        assert len(a) == len(b)
        assert np.all(np.abs(a - b) <= 1e-11*np.abs(a))
    gold = [-1.0062579942546936 + 1.3209787635623396e-05j, -1.004879704650626 + 0.0008946325671658642j]
    for r, g in zip(out[32][0], gold):
        assert len(r) == 1 and abs(r[0] - g) <= 1e-8*abs(g)
    assert len(out[32][0][2]) >= 1


def test_exit_predictor_against_literal_rule(ts):
    """The level loop's convergence predictor (csrc/psi_quad.cuh: stop once the last two increments show the
    quadratic collapse of the tanh-sinh rule and the extrapolated next one is 1e4 below the tolerance)
    against the reference's literal rule (tanhsinh.py:120-126) on every fixture point: M and D agree to far
    better than the 1e-10 of the parity contract, with fewer nodes."""
    bg = load("background.json")
    worst_m, worst_d, nodes = 0.0, 0.0, {"reference": 0, "predicted": 0}
    for case in load("disp.json")["cases"]:
        desc = product_desc(bg[case["dist"]]["spec"])
        w = carr(case["w"])
        out = {}
        for rule in ("reference", "predicted"):
            ctx = hs.HostsimContext(ts)
            ctx.exit_rule = rule
            did = ctx.add_dist(desc)
            out[rule] = ctx.disp_eval(did, case["kx"], case["kz"], case["alpha"], w, want_M=True)
            nodes[rule] += ctx.node_evals
        (d0, m0), (d1, m1) = out["reference"], out["predicted"]
        worst_m = max(worst_m, float(np.max(np.abs(m1 - m0)/np.max(np.abs(m0), axis=1, keepdims=True))))
        for i in range(len(w)):
            has_res = len(case["poles"][i]) > len(bg[case["dist"]]["poles"])
            rt = max(1e-11, resonant_floor(w[i]) if has_res else 0.0)
            scale = np.median(np.abs(d0))
            assert abs(d1[i] - d0[i]) <= rt*(abs(d0[i]) + scale), (case["dist"], w[i])
            worst_d = max(worst_d, abs(d1[i] - d0[i])/(abs(d0[i]) + scale))
    assert worst_m < 1e-11, worst_m
    assert nodes["predicted"] < 0.8*nodes["reference"], nodes
    print("predictor vs literal rule: worst M %.1e, worst D %.1e, nodes %d -> %d"
          % (worst_m, worst_d, nodes["reference"], nodes["predicted"]))


def test_fast_path_thread_logic(ts):
    """What one thread of K1's fast path computes (the one-lane instantiation of the quadrature with the
    loose predictor and a level cap -- exactly the code of disp_eval_thread_kernel): on every fixture point
    the result is either NaN ("not converged by level 8": the point goes to the warp-per-point kernel) or
    agrees with the literal rule to 1e-11, and with the cap lifted every point converges to that accuracy."""
    bg = load("background.json")
    n_fast = n_hard = nodes_fast = nodes_lit = 0
    for case in load("disp.json")["cases"]:
        desc = product_desc(bg[case["dist"]]["spec"])
        w = carr(case["w"])
        lit, fast, nocap = hs.HostsimContext(ts), hs.HostsimContext(ts), hs.HostsimContext(ts)
        fast.exit_rule = "fast"
        fast.set_level_cap(8)
        nocap.exit_rule = "fast"
        d0 = lit.disp_eval(lit.add_dist(desc), case["kx"], case["kz"], case["alpha"], w)
        d1 = fast.disp_eval(fast.add_dist(desc), case["kx"], case["kz"], case["alpha"], w)
        d2 = nocap.disp_eval(nocap.add_dist(desc), case["kx"], case["kz"], case["alpha"], w)
        scale = np.median(np.abs(d0))
        assert np.all(np.isfinite(d2))
        for i in range(len(w)):
            has_res = len(case["poles"][i]) > len(bg[case["dist"]]["poles"])
            rt = max(1e-11, resonant_floor(w[i]) if has_res else 0.0)
            assert abs(d2[i] - d0[i]) <= rt*(abs(d0[i]) + scale), (case["dist"], w[i])
            if np.isnan(d1[i].real):
                n_hard += 1
            else:
                n_fast += 1
                assert abs(d1[i] - d0[i]) <= rt*(abs(d0[i]) + scale), (case["dist"], w[i])
        nodes_fast += fast.node_evals
        nodes_lit += lit.node_evals
    assert nodes_fast < 0.25*nodes_lit
    assert n_fast > n_hard > 0, (n_fast, n_hard)
    print("fast path on the fixture points: %d converge within the cap, %d go to the warp kernel" % (n_fast, n_hard))
