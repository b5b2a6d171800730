"""The product's PSIMode / ClosedPath / scheduler classes end to end against fixtures produced by the
unmodified reference.  Every test runs on the CPU lane-emulation backend (conftest.hostsim_backend,
for the GPU-less container) and, under ``-m gpu``, through the real CUDA library."""
import numpy as np
import pytest

from common import carr, cplx, load, product_kwargs


def mode_cases():
    return load("mode.json")["cases"]


@pytest.mark.parametrize("case", mode_cases(), ids=lambda c: c["name"])
def test_psimode_reference_cases(backend, case):
    """test_psi_mode.py:32-189: same root (rtol 1e-7 as in the reference test, in fact ~1e-9) and
    same number of dispersion evaluations as the reference run."""
    from psitools_public_b200 import PSIMode, TanhSinh
    spec = load("background.json")[case["dist"]]["spec"]
    kw = product_kwargs(spec)
    kw.update(case["init"])
    np.random.seed(case["seed"])
    pm = PSIMode(tanhsinh_integrator=TanhSinh(), **kw)
    roots = pm.calculate(wave_number_x=case["K"][0], wave_number_z=case["K"][1],
                         viscous_alpha=case["alpha"])
    ref = carr(case["roots"])
    assert len(roots) == len(ref) == 1
    assert abs(roots[0] - ref[0]) <= 1e-8*abs(ref[0])
    pin = cplx(case["pin"])
    np.testing.assert_allclose(roots[0].real, pin.real, rtol=1e-7)
    np.testing.assert_allclose(roots[0].imag, pin.imag, rtol=1e-7)
    assert pm.n_function_call == case["n_function_call"]
    assert len(pm.z_sample) == len(case["z_sample"])
    np.testing.assert_array_equal(pm.z_sample, carr(case["z_sample"]))   # same random stream
    # the call was ONE search of the K4 pipeline (psi_mode_search_detail), not the host generators ...
    assert pm._device_search_ok([], False)
    # ... which leaves numpy's global stream exactly where the reference leaves it (2 numbers per sample point)
    after = np.random.random_sample()
    np.random.seed(case["seed"])
    np.random.random_sample(2*len(case["z_sample"]))
    assert after == np.random.random_sample()
    # ... and the attributes callers read: samples, support mask, a usable approximant
    assert pm.ra.maskF.shape == (len(pm.z_sample),) and 4 <= pm.ra.maskF.sum() <= len(pm.z_sample)/2 + 1
    assert abs(pm.ra.evaluate(roots[0])) < 1e-4*np.max(np.abs(pm.f_sample))
    # the generator form (engine='python': LAPACK AAA on the host) gives the same root and call count
    np.random.seed(case["seed"])
    pg = PSIMode(tanhsinh_integrator=TanhSinh(), engine="python", **kw)
    rg = pg.calculate(wave_number_x=case["K"][0], wave_number_z=case["K"][1], viscous_alpha=case["alpha"])
    assert len(rg) == 1 and abs(rg[0] - roots[0]) <= 1e-8*abs(roots[0])
    assert pg.n_function_call == case["n_function_call"]


@pytest.mark.parametrize("engine", ["device", "native", "python"])
def test_multistart_polish_tolerance(backend, engine):
    """multistart.json: the polish of a search with several start values of very different Im uses, for each start,
    the tolerance of THAT start (psi_mode.py:340-352 hands find_root one scalar at a time).  Cases 2 and 4 are the
    ones on which the product's approximant has the same zeros inside the domain as the reference's, so root sets
    and n_function_call must equal the reference's (30 and 124; the minimum over all starts costs more).  The other
    cases differ in spurious near-axis zeros of the approximant (the reference's D carries its LU noise) and are
    held to the root sets only."""
    import time
    from psitools_public_b200 import TanhSinh, psi_mode_mpi
    g = load("multistart.json")
    ts = TanhSinh()
    items = [{'__init__': {'stokes_range': g["stokes_range"], 'dust_to_gas_ratio': 3.0, 'real_range': [-2.0, 2.0],
                           'imag_range': [1e-8, 1.0], 'n_sample': 20, 'max_zoom_domains': 1, 'tol': 1.0e-13,
                           'clean_tol': 1e-4, 'tanhsinh_integrator': ts},
              'calculate': {'wave_number_x': c["kx"], 'wave_number_z': c["kz"]}, 'random_seed': g["seed"]}
             for c in g["cases"]]
    for k, (item, c) in enumerate(zip(items, g["cases"])):
        ms = psi_mode_mpi.MpiScheduler(time.time(), 1e9, verbose=False, engine=engine)
        res = ms.run_all([item])
        assert same_root_sets(list(res[0]), list(carr(c["roots"]))), (k, res[0])
        if k in (2, 4):
            assert ms.last_stats["function_calls"] == c["n_function_call"], (k, ms.last_stats["function_calls"])


def same_root_sets(got, ref, rtol=1e-8):
    got = sorted(got, key=lambda z: (z.real, z.imag))
    ref = sorted(ref, key=lambda z: (z.real, z.imag))
    if len(got) != len(ref):
        return False
    return all(abs(a - b) <= rtol*abs(b) for a, b in zip(got, ref))


def grid_arglist(n_axis, ts):
    ax = np.logspace(-1, 3, n_axis)
    kxg, kzg = np.meshgrid(ax, ax)
    return [{'__init__': {'stokes_range': [1e-8, 1.0], 'dust_to_gas_ratio': 3.0,
                          'size_distribution_power': 3.5, 'real_range': [-2.0, 2.0],
                          'imag_range': [1e-8, 1.0], 'n_sample': 20, 'max_zoom_domains': 1,
                          'tol': 1.0e-13, 'clean_tol': 1e-4, 'tanhsinh_integrator': ts},
             'calculate': {'wave_number_x': kx, 'wave_number_z': kz}, 'random_seed': 2}
            for kx, kz in zip(kxg.flatten(), kzg.flatten())]


@pytest.mark.parametrize("engine", ["device", "native", "python"])
@pytest.mark.parametrize("n_axis", [3, 6])
def test_mpischeduler_mode_grid(backend, n_axis, engine):
    """test_psi_mode_mpi.py:38-104 (3x3) and a 6x6 grid of the same family: root sets per pixel
    equal to the reference's (count exact, values to 1e-8 |omega|)."""
    import time
    from psitools_public_b200 import TanhSinhNoDeepCopy, psi_mode_mpi
    g = load("modegrid.json")["grid%d" % n_axis]
    ms = psi_mode_mpi.MpiScheduler(time.time(), 3600, verbose=False, engine=engine)
    res = ms.run(grid_arglist(n_axis, TanhSinhNoDeepCopy()))
    assert ms.last_stats["engine"] == engine
    assert sorted(res) == list(range(n_axis*n_axis))
    for i, case in enumerate(g["cases"]):
        assert same_root_sets(list(res[i]), list(carr(case["roots"]))), (i, res[i], case["roots"])
    if n_axis == 3:                              # goldens pasted in the reference test
        np.testing.assert_allclose(res[3].real, [-1.00487970465063e+00], rtol=1e-7)
        np.testing.assert_allclose(res[3].imag, [+8.94632567125043e-04], rtol=1e-7)
        np.testing.assert_allclose(res[6].real, [-1.00492431244162e+00], rtol=1e-7)
        np.testing.assert_allclose(res[6].imag, [+9.03855503901244e-04], rtol=1e-7)
    assert ms.last_stats["items"] == n_axis*n_axis and ms.last_stats["passes"] > 0


def test_dispersion_scheduler(backend):
    """test_psi_mode_mpi.py:108-171."""
    import time
    from psitools_public_b200 import psi_mode_mpi
    items = [{'__init__': {'stokes_range': [1e-8, 1.0], 'dust_to_gas_ratio': 3.0,
                           'real_range': [-2.0, 2.0], 'imag_range': [-2.0, 2.0]},
              'dispersion': {'w': re + 1j*im, 'wave_number_x': 1e2, 'wave_number_z': 1e1,
                             'viscous_alpha': 0.0}}
             for im in (-2.0, 2.0) for re in (-2.0, 2.0)]
    res = psi_mode_mpi.DispersionRelationMpiScheduler(time.time(), 3600, verbose=False).run(items)
    np.testing.assert_allclose(res[0].real, -1.72733464670814e+08, rtol=1e-7)
    np.testing.assert_allclose(res[0].imag, +5.32484293672187e+07, rtol=1e-7)
    np.testing.assert_allclose(res[3].real, -8.18797751107521e+07, rtol=1e-7)
    np.testing.assert_allclose(res[3].imag, +8.84518043425989e+07, rtol=1e-7)
    assert res[1].shape == ()
    # array-valued 'w' items (the reference passes w straight to PSIMode.dispersion, which takes any shape)
    w2 = np.array([[-2 - 2j, 2 - 2j], [-2 + 2j, 2 + 2j]])
    mixed = [dict(items[0], dispersion=dict(items[0]['dispersion'], w=w2)), items[3],
             dict(items[0], dispersion=dict(items[0]['dispersion'], w=w2.ravel()[:3]))]
    res2 = psi_mode_mpi.DispersionRelationMpiScheduler(time.time(), 3600, verbose=False).run(mixed)
    assert res2[0].shape == (2, 2) and res2[1].shape == () and res2[2].shape == (3,)
    flat = np.array([res[i] for i in range(4)])
    np.testing.assert_allclose(res2[0].ravel(), flat, rtol=1e-13)
    np.testing.assert_allclose(res2[2], flat[:3], rtol=1e-13)
    np.testing.assert_allclose(res2[1], res[3], rtol=1e-13)


def count_arglist(ts):
    z_list = [-2.0 + 2e-7j, 2.0 + 2e-7j, 2.0 + 2.0j, -2.0 + 2.0j]
    ax = np.logspace(-1, 3, 3)
    kxg, kzg = np.meshgrid(ax, ax)
    return [{'PSIDispersion.__init__': {'stokes_range': [1e-8, 1.0], 'dust_to_gas_ratio': 3.0,
                                        'size_distribution_power': 3.5, 'tanhsinh_integrator': ts},
             'dispersion': {'wave_number_x': kx, 'wave_number_z': kz, 'viscous_alpha': 0.0},
             'z_list': z_list, 'count_roots': {'tol': 0.1}}
            for kx, kz in zip(kxg.flatten(), kzg.flatten())]


@pytest.mark.parametrize("engine", ["device", "python"])
def test_root_count_map(backend, engine):
    """test_complex_roots_mpi.py:32-102: counts [0,0,0,1,0,0,1,0,0] exactly."""
    import time
    from psitools_public_b200 import TanhSinhNoDeepCopy, complex_roots_mpi
    res = complex_roots_mpi.MpiScheduler(time.time(), 3600, verbose=False,
                                         engine=engine).run(
        count_arglist(TanhSinhNoDeepCopy()))
    assert [res[i] for i in range(9)] == [0, 0, 0, 1, 0, 0, 1, 0, 0]
    gold = load("contour.json")["cases"]
    assert [c["count"] for c in gold] == [res[i] for i in range(9)]


def test_closedpath_dropin(backend):
    """ClosedPath driven by an arbitrary callable: same refinement history as the reference
    (number of points added per pass) on the contour with one root."""
    from psitools_public_b200 import ClosedPath, PSIDispersion, TanhSinh
    gold = load("contour.json")["cases"][3]
    pd = PSIDispersion(3.0, [1e-8, 1.0], tanhsinh_integrator=TanhSinh())
    cp = ClosedPath(lambda z: pd.calculate(z, gold["kx"], gold["kz"]),
                    [-2.0 + 2e-7j, 2.0 + 2e-7j, 2.0 + 2.0j, -2.0 + 2.0j])
    assert cp.count_roots(tol=0.1) == gold["count"] == 1
    assert cp.success
    # the refinement is driven by arg(D) jumps > 0.1: the point count follows the reference closely
    assert abs(len(cp.points) - gold["n_points"]) <= 0.05*gold["n_points"]
    with pytest.raises(RuntimeError):
        ClosedPath(lambda z: pd.calculate(z, gold["kx"], gold["kz"]),
                   [-2.0 + 2e-7j, 2.0 + 2e-7j, 2.0 + 2.0j, -2.0 + 2.0j]).count_roots(max_iter=3)


def test_mode_with_count_and_guess(backend):
    """count_roots=True joins the contour points to the samples (psi_mode.py:127-138); a guessed
    root adds a zoom domain (psi_mode.py:141-145)."""
    from psitools_public_b200 import PSIMode, TanhSinh
    np.random.seed(0)
    pm = PSIMode(3.0, [1e-8, 1.0], [-2, 2], [1e-8, 1], tanhsinh_integrator=TanhSinh())
    roots = pm.calculate(0.1, 10.0, guess_roots=[-1.0048797 + 8.9463e-4j])
    assert len(roots) >= 1
    assert min(abs(r - (-1.004879704650626 + 0.0008946325671658642j)) for r in roots) < 1e-9
    assert pm.n_function_call >= 40


def test_polish_kernel_k3(backend):
    """Kernel K3 (two-stage secant polish on the device) against the Python statement of the same
    logic (psi_mode.polish_machine) driven through the lock-step engine: same roots, same sentinel
    pattern, same number of dispersion evaluations."""
    from psitools_public_b200 import PSIMode, TanhSinh
    from psitools_public_b200._engine import Job, run_lockstep
    from psitools_public_b200.psi_mode import ModeTrace, polish_machine
    pm = PSIMode(3.0, [1e-8, 1.0], [-2, 2], [1e-8, 1], tanhsinh_integrator=TanhSinh())
    ctx, did = pm.psi_dispersion.ctx, pm.psi_dispersion.dist_id
    root = -1.004879704650626 + 0.0008946325671658642j
    starts = [[root*(1 + 1e-4), 0.5 + 0.3j, -1.2 + 0.05j], [root + 1e-5], [1.5 + 0.9j]]
    ks = [(0.1, 10.0), (0.1, 10.0), (0.1, 10.0)]
    max_iter = [20, 100, 5]
    got, ncalls, mconv, status = ctx.polish_batch([did]*3, [k[0] for k in ks], [k[1] for k in ks], [0.0]*3,
                                                  [[-2, 2, 1e-8, 1]]*3, starts, max_iter)
    assert list(status) == [0, 0, 0]
    for i in range(3):
        tr = ModeTrace()
        pm._sync_cfg()
        job = Job(polish_machine(pm.cfg, starts[i], max_iter[i], tr), did, ks[i][0], ks[i][1], 0.0)
        run_lockstep(ctx, [job])
        kept = [z for z in got[i] if pm.domain.is_in(z)]
        assert same_root_sets(kept, list(job.result)), (i, got[i], job.result)
        assert ncalls[i] == tr.n_function_call, (i, ncalls[i], tr.n_function_call)
    assert abs(got[0][0] - root) < 1e-9 and abs(got[1][0] - root) < 1e-9
    assert mconv[1] >= 1
    # a start value on the real axis makes the reference's tolerance zero -> ValueError there
    _, _, _, st = ctx.polish_batch([did], [0.1], [10.0], [0.0], [[-2, 2, 1e-8, 1]], [[0.3 + 0j]], [20])
    assert st[0] == 1


def test_contour_kernel_k2_status(backend):
    """Kernel K2 error paths: max_iter exhausted -> RuntimeError in the reference -> count -666."""
    from psitools_public_b200 import PSIDispersion, TanhSinh
    pd = PSIDispersion(3.0, [1e-8, 1.0], tanhsinh_integrator=TanhSinh())
    z = [-2.0 + 2e-7j, 2.0 + 2e-7j, 2.0 + 2.0j, -2.0 + 2.0j]
    counts, status, npts, st = pd.ctx.contour_batch([pd.dist_id]*2, [0.1, 0.1], [10.0, 10.0], [0.0, 0.0],
                                                    [z, z], [0.1, 0.1], [0.1, 0.1], [3, 100])
    assert counts[0] == -666 and status[0] == 2
    assert counts[1] == 1 and status[1] == 0
    gold = load("contour.json")["cases"][3]
    assert abs(npts[1] - gold["n_points"]) <= 0.05*gold["n_points"]
    assert st["evals"] >= npts[1]


def test_mode_count_roots_true(backend):
    """PSIMode.calculate(count_roots=True) (psi_mode.py:127-138): the contour count is stored in
    n_roots and the contour points join the AAA samples."""
    from psitools_public_b200 import PSIMode, TanhSinh
    np.random.seed(0)
    pm = PSIMode(3.0, [1e-8, 1.0], [-2, 2], [2e-7, 2], tanhsinh_integrator=TanhSinh())
    roots = pm.calculate(0.1, 10.0, count_roots=True)
    assert pm.n_roots == 1                                   # same contour as test_complex_roots_mpi pixel 3
    assert len(pm.z_sample) > 200                            # 20 random samples + the contour points
    assert min(abs(r - (-1.004879704650626 + 0.0008946325671658642j)) for r in roots) < 1e-8


def test_mode_dispersion_map_and_minimum(backend):
    """plot_dispersion's arrays (psi_mode.py:507-517): the exact D of the whole plot grid comes from
    one batched evaluation and equals per-frequency calls; the rational approximation reproduces it in
    the domain; find_minimum (psi_mode.py:399-409) returns a point of the domain below every grid value."""
    from psitools_public_b200 import PSIMode, TanhSinh
    np.random.seed(0)
    pm = PSIMode(3.0, [1e-8, 0.01], [-2, 2], [1e-8, 1], tanhsinh_integrator=TanhSinh())
    roots = pm.calculate(10, 1000)
    assert len(roots) == 1
    x, y, f_appro, f_exact = pm.dispersion_map(10, 1000, N=6, show_exact=True)
    assert f_appro.shape == f_exact.shape == (6, 6) and x[0] == -2 and y[-1] == 1
    for (i, j) in [(0, 0), (2, 3), (5, 5)]:
        one = pm.dispersion(x[j] + 1j*y[i], 10, 1000)
        assert abs(f_exact[i, j] - one) <= 1e-12*abs(one)
    inner = f_exact[1:, :]                                      # rows with Im w > 1e-8 (off the branch cut)
    assert np.max(np.abs(f_appro[1:, :] - inner)/np.abs(inner)) < 1e-6
    _, _, fa, fe = pm.dispersion_map(10, 1000, x=np.array([-1.0, 0.5]), y=np.array([0.1, 0.2]))
    assert fa.shape == (2, 2) and fe is None
    zmin = pm.find_minimum()                                    # a root, or the trace of one just outside
    assert pm.domain.xmin <= zmin.real <= pm.domain.xmax and pm.domain.ymin <= zmin.imag <= pm.domain.ymax
    assert abs(pm.ra.evaluate(zmin)) <= np.min(np.abs(f_appro))
    try:
        import matplotlib  # noqa: F401
    except ImportError:
        with pytest.raises(ImportError):
            pm.plot_dispersion(10, 1000, N=4)
