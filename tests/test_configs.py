"""Small-scale parity for the remaining BASELINE configurations.

config 5: parameter sweep -- many (mu, tau_max) dust distributions x a (Kx, Kz) grid of PSIMode solves
          in ONE batch (mixed distribution ids per item), incl. the high-order table TanhSinh(15, 13)
          whose 408 KB node table no longer fits shared memory (K1 reads it through L1/L2 instead).
config 4: covered by test_grid_refine.py.  config 2/3: test_disp_parity.py / test_host_logic.py."""
import time

import numpy as np
import pytest

from test_grid_refine import distinct
from test_host_logic import same_root_sets


def sweep_items(ts, mus, taumaxs, n_axis):
    ax = np.logspace(-1, 3, n_axis)
    items = []
    for mu in mus:
        for tmax in taumaxs:
            for kx in ax:
                for kz in ax:
                    items.append({'__init__': {'stokes_range': [1e-8, tmax], 'dust_to_gas_ratio': mu,
                                               'real_range': [-2.0, 2.0], 'imag_range': [1e-8, 1.0],
                                               'n_sample': 20, 'max_zoom_domains': 1,
                                               'tanhsinh_integrator': ts},
                                  'calculate': {'wave_number_x': kx, 'wave_number_z': kz},
                                  'random_seed': 2})
    return items


def oracle_roots(item, tab):
    from oracle import psi_oracle as po
    init = {k: v for k, v in item['__init__'].items() if k != 'tanhsinh_integrator'}
    return po.run_mode_item({'__init__': init, 'calculate': item['calculate'],
                             'random_seed': item['random_seed']}, tab)


def test_config5_parameter_sweep(backend):
    from oracle import psi_oracle as po
    from psitools_public_b200 import TanhSinhNoDeepCopy, psi_mode_mpi
    ts = TanhSinhNoDeepCopy()
    items = sweep_items(ts, [0.1, 10.0], [1e-3, 1.0], 3)          # 4 distributions x 3x3 K
    ms = psi_mode_mpi.MpiScheduler(time.time(), 1e9, verbose=False)
    res = ms.run(items)
    assert ms.last_stats["engine"] == "device" and len(res) == 36
    tab = po.NodeTable()
    n_with_roots = 0
    for i in range(0, 36, 2):                                     # every other item against the oracle
        ref = oracle_roots(items[i], tab)
        # root sets are compared de-duplicated: how often the searches of one pixel land on the same
        # root is an internal-trajectory quantity, the set of roots is not
        assert same_root_sets(distinct(res[i]), distinct(ref)), (i, items[i]['calculate'], res[i], ref)
        n_with_roots += len(ref) > 0
    assert n_with_roots >= 2


def test_high_order_table_level13(backend):
    """TanhSinh(15, 13): 25,485 nodes; D against the oracle run with the same table."""
    from oracle import psi_oracle as po
    from psitools_public_b200 import PSIDispersion, TanhSinh
    ts = TanhSinh(15, 13)
    assert len(ts.xj) == 25485
    pd = PSIDispersion(3.0, [1e-8, 1.0], tanhsinh_integrator=ts)
    w = np.array([0.3 + 0.2j, -1.0048 + 1e-4j, 1.5 - 0.7j, -0.2 + 1e-6j])
    got, M = pd.calculate(w, 10.0, 100.0, return_M=True)
    od = po.Dispersion(po.DustDist.power_law(3.0, [1e-8, 1.0]), po.NodeTable(15, 13))
    ref, Mref = od.calculate(w, 10.0, 100.0, return_M=True)
    assert abs(pd.vgx - od.vgx) < 1e-14*abs(od.vgx)
    for i in range(len(w)):
        assert np.max(np.abs(M[i] - Mref[i])) <= 2e-9*np.max(np.abs(Mref[i]))
        assert abs(got[i] - ref[i]) <= 2e-9*abs(ref[i])
    # levels 12 and 13 agree (the extra level only confirms convergence)
    pd12 = PSIDispersion(3.0, [1e-8, 1.0], tanhsinh_integrator=TanhSinh(15, 12))
    np.testing.assert_allclose(pd12.calculate(w[:3], 10.0, 100.0), got[:3], rtol=1e-11)


def test_ragged_batch_and_empty(backend):
    """One batch with different n_sample / max_zoom_domains / guess lists / seeds / distributions per
    item: the device pipeline (engine='device'), the LAPACK engine ('native') and the Python generators
    ('python') return the same de-duplicated root sets; an empty arg list gives an empty dict."""
    from psitools_public_b200 import TanhSinhNoDeepCopy, psi_mode_mpi
    ts = TanhSinhNoDeepCopy()
    root = -1.004879704650626 + 0.0008946325671658642j

    def item(tmax, kx, kz, n_sample, max_zoom, seed, guess=()):
        return {'__init__': {'stokes_range': [1e-8, tmax], 'dust_to_gas_ratio': 3.0,
                             'real_range': [-2.0, 2.0], 'imag_range': [1e-8, 1.0], 'n_sample': n_sample,
                             'max_zoom_domains': max_zoom, 'tanhsinh_integrator': ts},
                'calculate': {'wave_number_x': kx, 'wave_number_z': kz, 'guess_roots': list(guess)},
                'random_seed': seed}

    items = [item(1.0, 0.1, 10.0, 20, 1, 2), item(1.0, 0.1, 10.0, 15, 0, 5, [root*(1 + 1e-3)]),
             item(1e-2, 10.0, 1000.0, 20, 1, 0), item(1.0, 1.0, 1.0, 12, 2, 7),
             item(1.0, 0.1, 10.0, 20, 0, 3, [root, root + 1e-4, 0.3 + 0.1j])]
    res = {}
    for eng in ("device", "native", "python"):
        ms = psi_mode_mpi.MpiScheduler(time.time(), 1e9, verbose=False, engine=eng)
        assert ms.run([]) == {}
        res[eng] = ms.run(items)
        assert ms.last_stats["engine"] == eng
    for i in range(len(items)):
        a, b, c = (distinct(res[e][i]) for e in ("device", "native", "python"))
        assert same_root_sets(a, c) and same_root_sets(b, c), (i, res["device"][i], res["native"][i], res["python"][i])
    assert any(abs(r - root) < 1e-9 for r in res["device"][0])
    assert any(abs(r - root) < 1e-9 for r in res["device"][4])
    assert abs(res["device"][2][0] - (-1.0062579942546936 + 1.3209787635623396e-05j)) < 1e-7


@pytest.mark.parametrize("engine", ["device", "native"])
def test_config3_strided_sample_vs_reference(backend, engine):
    """BASELINE config 3 parity sample: the 16x16 strided subset of the 256x256 (Kx,Kz) map against the
    UNMODIFIED reference (fixture modegrid16.json): de-duplicated root sets per pixel must be equal --
    the root count exactly, the values to 1e-8 |omega|.  (The CPU backend checks every fourth pixel to
    keep the suite short; the GPU run checks all 256.)"""
    from common import carr, load
    from psitools_public_b200 import TanhSinhNoDeepCopy, psi_mode_mpi
    g = load("modegrid16.json")
    cases = g["cases"] if backend == "gpu" else g["cases"][::4]
    ts = TanhSinhNoDeepCopy()
    items = [{'__init__': {'stokes_range': [1e-8, 1.0], 'dust_to_gas_ratio': 3.0,
                           'size_distribution_power': 3.5, 'real_range': [-2.0, 2.0],
                           'imag_range': [1e-8, 1.0], 'n_sample': 20, 'max_zoom_domains': 1,
                           'tol': 1.0e-13, 'clean_tol': 1e-4, 'tanhsinh_integrator': ts},
              'calculate': {'wave_number_x': c["kx"], 'wave_number_z': c["kz"]}, 'random_seed': 2}
             for c in cases]
    res = psi_mode_mpi.MpiScheduler(time.time(), 1e9, verbose=False, engine=engine).run(items)
    bad = [i for i, c in enumerate(cases)
           if not same_root_sets(distinct(res[i]), distinct(carr(c["roots"])))]
    n_ref_roots = sum(len(distinct(carr(c["roots"]))) for c in cases)
    n_roots = sum(len(distinct(res[i])) for i in range(len(cases)))
    print("pixels %d, with roots (reference) %d, distinct roots %d (reference %d), mismatching pixels %s"
          % (len(cases), sum(len(c["roots"]) > 0 for c in cases), n_roots, n_ref_roots, bad))
    assert n_roots == n_ref_roots
    assert not bad, [(cases[i]["kx"], cases[i]["kz"], res[i], cases[i]["roots"]) for i in bad]


@pytest.mark.gpu
def test_config3_full_size_properties(built):
    """Size-independent properties at a large size (128x128 map, 16,384 searches, device engine):
    (1) every reported root is a root of the exact dispersion relation: |D(w)| << |D(w + 1e-6) - D(w)|
        (a first-order bound |dw| < 1e-9 on the distance to the true root), checked in ONE K1 batch;
    (2) the argument principle agrees: on a 24x24 sub-map the number of distinct roots found inside the
        search rectangle equals the winding number of D around that rectangle (kernel K2) wherever a
        root was found, and pixels where the contour count is zero have no root;
    (3) a second run returns bit-identical roots (deterministic kernels)."""
    from psitools_public_b200 import PSIMode, TanhSinhNoDeepCopy
    ts = TanhSinhNoDeepCopy()
    pm = PSIMode(3.0, [1e-8, 1.0], [-2, 2], [1e-8, 1.0], tanhsinh_integrator=ts)
    ctx, did = pm.psi_dispersion.ctx, pm.psi_dispersion.dist_id
    n = 128
    ax = np.logspace(-1, 3, n)
    kxg, kzg = np.meshgrid(ax, ax)
    kx, kz, N = kxg.ravel(), kzg.ravel(), n*n
    dom = [[-2, 2, 1e-8, 1.0]]*N
    run = lambda: ctx.mode_batch([did]*N, kx, kz, np.zeros(N), dom, 20, 1, 100, 1e-13, 1e-4, [2]*N, [[]]*N,
                                 device_aaa=True)
    roots, ncalls, st = run()
    roots2, _, _ = run()
    assert all(np.array_equal(a, b) for a, b in zip(roots, roots2))                      # (3)
    idx = np.array([i for i, r in enumerate(roots) for _ in r], dtype=int)
    w = np.array([z for r in roots for z in r])
    assert len(w) > 2000 and np.all(w.imag > 0)
    d0 = ctx.disp_eval(np.full(len(w), did, np.int32), kx[idx], kz[idx], np.zeros(len(w)), w)
    d1 = ctx.disp_eval(np.full(len(w), did, np.int32), kx[idx], kz[idx], np.zeros(len(w)), w + 1e-6)
    assert np.all(np.abs(d0) < 1e-3*np.abs(d1 - d0)), np.max(np.abs(d0)/np.abs(d1 - d0))    # (1)
    # (2) on a 24x24 strided sub-map
    sub = [i*n + j for i in range(2, n, 5) for j in range(2, n, 5)][:576]
    z_list = [-2 + 1e-8j, 2 + 1e-8j, 2 + 1j, -2 + 1j]
    counts, status, _, _ = ctx.contour_batch([did]*len(sub), kx[sub], kz[sub], np.zeros(len(sub)),
                                             [z_list]*len(sub), 0.1, 0.1, 100)
    assert np.all(status == 0)
    for c, i in zip(counts, sub):
        found = len(distinct(roots[i]))
        if found:
            assert found == c, (kx[i], kz[i], roots[i], c)
        if c == 0:
            assert found == 0


@pytest.mark.gpu
def test_device_vs_lapack_engine_other_distributions(built):
    """K4's AAA (Householder QR + inverse iteration, Ehrlich-Aberth) against the LAPACK engine on
    distributions other than the MRN map of the fixtures: PowerBumpTail (48x48 map) and a viscous MRN case.
    Root sets must agree on at least 99.5 % of the pixels, and wherever they differ the two sets may differ
    by one root only, the rest equal to 1e-8 (measured on larger maps: 0.007 % of pixels differ, each time
    by one root that one rational approximation resolves and the other does not, see
    profiles/engine_parity_r1c.txt)."""
    from psitools_public_b200 import PowerBumpTail, PSIMode, SizeDensity, TanhSinhNoDeepCopy
    ts = TanhSinhNoDeepCopy()
    pb = PowerBumpTail(amin=1e-8, aP=0.1, aL=2.0/3.0*0.1, aR=0.156, bumpfac=2.0)
    sd = SizeDensity(pb.sigma0, [1e-8, 0.156])
    sd.poles = [pb.get_discontinuity()]
    cases = [(PSIMode(10.0, [1e-8, 0.156], [-2, 2], [1e-7, 1.0], size_density=sd, tanhsinh_integrator=ts), 0.0),
             (PSIMode(3.0, [1e-8, 0.1], [-2, 2], [1e-8, 1.0], tanhsinh_integrator=ts), 1e-6)]
    n = 48
    ax = np.logspace(-1, 3, n)
    kxg, kzg = np.meshgrid(ax, ax)
    N = n*n
    for pm, alpha in cases:
        ctx, did = pm.psi_dispersion.ctx, pm.psi_dispersion.dist_id
        dom = [[pm.domain.xmin, pm.domain.xmax, pm.domain.ymin, pm.domain.ymax]]*N
        out = {}
        for dev in (True, False):
            out[dev], _, _ = ctx.mode_batch([did]*N, kxg.ravel(), kzg.ravel(), np.full(N, alpha), dom, 20, 1,
                                            100, 1e-13, 1e-4, [2]*N, [[]]*N, device_aaa=dev)
        differ = 0
        for i in range(N):
            a, b = distinct(out[True][i]), distinct(out[False][i])
            if not same_root_sets(a, b):
                differ += 1
                small, big = (a, b) if len(a) <= len(b) else (b, a)
                assert len(big) - len(small) <= 1, (i, a, b)
                assert all(any(abs(x - y) <= 1e-8*abs(y) for x in big) for y in small), (i, a, b)
        assert differ <= 0.005*N, differ


def test_mode_batch_chunking_is_invisible(backend, monkeypatch):
    """Ctx.mode_batch cuts batches that would not fit the device budget into pieces; items are independent,
    so the result must not depend on the cut (here: a budget that forces three pieces, with guess lists
    in CSR form, against the uncut call)."""
    from psitools_public_b200 import PSIMode, TanhSinhNoDeepCopy
    pm = PSIMode(3.0, [1e-8, 1e-2], [-2, 2], [1e-8, 1.0], tanhsinh_integrator=TanhSinhNoDeepCopy())
    ctx, did = pm.psi_dispersion.ctx, pm.psi_dispersion.dist_id
    kx = np.array([10.0, 10.0, 30.0, 10.0, 3.0])
    kz = np.array([1000.0, 100.0, 300.0, 1000.0, 30.0])
    guesses = (np.array([-1.0 + 1e-5j, 0.5 + 0.1j, -1.0 + 2e-5j]), np.array([0, 1, 1, 3, 3, 3]))
    call = lambda: ctx.mode_batch([did]*5, kx, kz, np.zeros(5), [[-2, 2, 1e-8, 1.0]]*5, 20, 1, 100, 1e-13,
                                  1e-4, [0, 1, 2, 0, 1], guesses, device_aaa=True, dense=True)
    (r0, n0), c0, s0 = call()
    monkeypatch.setenv("PSI_B200_BATCH_BYTES", "17000")
    (r1, n1), c1, s1 = call()
    assert s1.get("chunks") == 3 and "chunks" not in s0
    assert np.array_equal(n0, n1) and np.array_equal(c0, c1) and n0.sum() >= 2
    assert all(np.array_equal(r0[i, :n0[i]], r1[i, :n1[i]]) for i in range(5))
    assert s0["evals"] == s1["evals"]


def test_root_slot_overflow_redoes_only_those_searches(backend):
    """More roots than result slots (a refinement pixel whose guesses lead to the same root returns it several
    times): only the overflowing searches are run again with room, the answer equals the one obtained with enough
    slots from the start.  The searches are those of the tiny refinement campaign of test_grid_refine.py."""
    from test_grid_refine import make_refiner
    gr, log = make_refiner("device")
    if backend == "gpu":                     # a campaign large enough to contain such searches (35 x 35 map)
        gr.nbase, gr.krange, gr.reruns = (16, 16), (-1.0, 3.0), 1
    gr.run_basegrid()
    gr.fill_in_grid()
    items = [(kx, kz, list(g)) for batch in log[1:] for kx, kz, g, _ in batch]
    if max(len(r) for batch in log[1:] for _, _, _, r in batch) < 2:
        assert backend != "gpu"
        pytest.skip("no search with two roots in the tiny CPU campaign")
    pm = gr.ms._cache.get(gr._shared_init(0))
    ctx, did = pm.psi_dispersion.ctx, pm.psi_dispersion.dist_id
    n = len(items)
    args = ([did]*n, [i[0] for i in items], [i[1] for i in items], np.zeros(n), [[-2, 2, 1e-7, 1.0]]*n, 20, 0,
            100, 1e-13, 1e-4, [2]*n, [i[2] for i in items])
    (r8, n8), c8, _ = ctx.mode_batch(*args, max_roots=8, device_aaa=True, dense=True)
    (r1, n1), c1, _ = ctx.mode_batch(*args, max_roots=1, device_aaa=True, dense=True)
    assert n8.max() >= 2 and r1.shape[1] == n8.max()
    assert np.array_equal(n1, n8) and np.array_equal(c1, c8)
    for i in range(n):
        assert np.array_equal(r1[i, :n1[i]], r8[i, :n8[i]]), i


@pytest.mark.gpu
def test_k3_block_mode_equals_warp_mode(built, monkeypatch):
    """K3 polishes a launch with few root searches with one 384-thread block per search (CtaCuda policy)
    instead of one warp: same code, different summation order inside each D evaluation.  On a 24x24 map the
    two modes must find the same roots (1e-11 relative -- the iteration stops at |step| < 1e-8..1e-13, so
    last-bit differences in D move the converged value by far less) with the same number of D
    evaluations on all but a few searches, and both must match the LAPACK engine."""
    from psitools_public_b200 import PSIMode, TanhSinhNoDeepCopy
    pm = PSIMode(3.0, [1e-8, 1.0], [-2, 2], [1e-8, 1.0], tanhsinh_integrator=TanhSinhNoDeepCopy())
    ctx, did = pm.psi_dispersion.ctx, pm.psi_dispersion.dist_id
    n = 24
    ax = np.logspace(-1, 3, n)
    kxg, kzg = np.meshgrid(ax, ax)
    N = n*n
    run = lambda dev=True: ctx.mode_batch([did]*N, kxg.ravel(), kzg.ravel(), np.zeros(N), [[-2, 2, 1e-8, 1.0]]*N, 20, 1,
                                          100, 1e-13, 1e-4, [2]*N, [[]]*N, device_aaa=dev)
    monkeypatch.setenv("PSI_K3_CTA_ITEMS", "0")
    warp, calls_w, _ = run()
    monkeypatch.setenv("PSI_K3_CTA_ITEMS", "4096")
    block, calls_b, _ = run()
    lapack, _, _ = run(False)
    assert sum(len(r) > 0 for r in warp) > 50
    for a, b, c in zip(warp, block, lapack):
        assert len(a) == len(b)
        assert np.allclose(a, b, rtol=1e-11, atol=0)
        assert same_root_sets(distinct(b), distinct(c))
    assert np.count_nonzero(calls_w != calls_b) <= 0.02*N
