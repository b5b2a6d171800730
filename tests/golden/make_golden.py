"""Generate golden fixtures by running the UNMODIFIED reference in the build container.

    python tests/golden/make_golden.py [section ...]      (sections: see SECTIONS at the bottom)

The reference (pure Python, /root/reference) cannot travel to the GPU box, so its outputs on
fixed inputs are committed as small JSON files next to this script.  Complex numbers are stored
as [re, im] pairs of repr-exact doubles.  Only the import shim of refshim.py is applied; no
arithmetic of the reference is touched.
"""
import hashlib
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import refshim  # noqa: E402

PROCS = int(os.environ.get("GOLDEN_PROCS", "8"))


def cj(z):
    z = complex(z)
    return [z.real, z.imag]


def cl(zs):
    return [cj(z) for z in np.ravel(zs)]


def dump(name, obj):
    path = os.path.join(HERE, name)
    with open(path, "w") as fh:
        json.dump(obj, fh, indent=0)
    print("wrote", path, os.path.getsize(path), "bytes", flush=True)


# ------------------------------------------------------------------------------------------
# distributions used throughout (built with the reference's own classes)
# ------------------------------------------------------------------------------------------
DISTS = {
    "mrn_1e-2": dict(kind="powerlaw", mu=3.0, stokes_range=[1e-8, 1e-2]),
    "mrn_1": dict(kind="powerlaw", mu=3.0, stokes_range=[1e-8, 1.0]),
    "mrn_1e-1": dict(kind="powerlaw", mu=3.0, stokes_range=[1e-8, 1e-1]),
    "pl32_c10": dict(kind="powerlaw", mu=0.5, stokes_range=[1e-6, 0.5], power=3.2, c=10.0),
    "bump": dict(kind="powerbump", mu=10.0, amin=1e-8, aP=0.1, aL=2.0/3.0*0.1, aR=0.156,
                 bumpfac=2.0),
    "bumptail": dict(kind="powerbumptail", mu=10.0, amin=1e-8, aP=0.1, aL=2.0/3.0*0.1, aR=0.156,
                     bumpfac=2.0, aBT=1e-6),
    "bump_eps": dict(kind="powerbump", mu=2.0, amin=1e-7, aP=0.05, aL=None, aR=None,
                     bumpfac=3.0, epstot=2.0),
    "single": dict(kind="powerlaw", mu=3.0, stokes_range=[1e-8, 0.1], single=True),
}


def ref_init_kwargs(ns, spec, with_ts=True):
    """PSIDispersion kwargs for a DISTS entry, using reference classes (test_psi_mode.py:89-105)."""
    kw = dict(dust_to_gas_ratio=spec["mu"])
    if spec["kind"] == "powerlaw":
        kw["stokes_range"] = spec["stokes_range"]
        kw["size_distribution_power"] = spec.get("power", 3.5)
        kw["single_size_flag"] = spec.get("single", False)
    else:
        if spec["kind"] == "powerbump":
            pb = ns.power_bump.PowerBump(amin=spec["amin"], aP=spec["aP"], aL=spec["aL"],
                                         aR=spec["aR"], bumpfac=spec["bumpfac"],
                                         epstot=spec.get("epstot"))
        else:
            pb = ns.power_bump.PowerBumpTail(amin=spec["amin"], aBT=spec["aBT"], aP=spec["aP"],
                                             aL=spec["aL"], aR=spec["aR"],
                                             bumpfac=spec["bumpfac"], epstot=spec.get("epstot"))
        sd = ns.sizedensity.SizeDensity(pb.sigma0, [spec["amin"], pb.aR])
        sd.poles = [pb.get_discontinuity()]
        kw["stokes_range"] = [spec["amin"], pb.aR]
        kw["size_density"] = sd
    kw["sound_speed_over_eta"] = spec.get("c", 20.0)
    if with_ts:
        kw["tanhsinh_integrator"] = ns.tanhsinh.TanhSinh()
    return kw


# ------------------------------------------------------------------------------------------
def sec_nodes():
    ns = refshim.load()
    out = {}
    for p, m in ((15, 12), (15, 10), (15, 13), (12, 8)):
        ts = ns.tanhsinh.TanhSinh(p, m)
        out["%d_%d" % (p, m)] = dict(
            n=len(ts.xj),
            sha_x=hashlib.sha256(np.ascontiguousarray(ts.xj).tobytes()).hexdigest(),
            sha_w=hashlib.sha256(np.ascontiguousarray(ts.wj).tobytes()).hexdigest(),
            x_every_499=list(map(float, ts.xj[::499])), w_every_499=list(map(float, ts.wj[::499])),
            x_last=float(ts.xj[-1]), w_last=float(ts.wj[-1]))
    # known-answer integrals of psitools_examples/tanhsinh_example.ipynb
    ts = ns.tanhsinh.TanhSinh()
    eps = 1e-8
    known = {
        "xlog1px": float(ts.integrate(lambda x: x*np.log(1 + x), 0.0, 1.0)),
        "sqrtxlogx": float(ts.integrate(lambda x: np.sqrt(x)*np.log(x), 0.0, 1.0)),
        "log2x": float(ts.integrate(lambda x: np.log(x)**2, 0.0, 1.0)),
        "lorentz": float(ts.integrate(lambda x: eps/(x*x + eps*eps), 0.0, 1.0)),
        "cexp": cj(ts.integrate(lambda x: np.exp((1 + 2j)*x), -1.0, 2.5)),
    }
    out["known"] = known
    dump("nodes.json", out)


def _background_one(name):
    ns = refshim.load()
    spec = DISTS[name]
    pd = ns.psi_dispersion.PSIDispersion(**ref_init_kwargs(ns, spec))
    s = np.logspace(np.log10(pd.taumin/pd.taumax), 0, 23)
    return name, dict(spec=spec, vgx=float(pd.vgx), vgy=float(pd.vgy),
                      base_poly=list(map(float, pd.base_poly)),
                      rhod=float(pd.size_density.rhod),
                      poles=list(map(float, pd.size_density.poles)),
                      s_probe=list(map(float, s)),
                      F_probe=list(map(float, np.asarray(pd.size_density(s), dtype=float))),
                      ux_probe=list(map(float, pd.ux(s*pd.taumax))),
                      uy_probe=list(map(float, pd.uy(s*pd.taumax))))


def sec_background():
    import multiprocessing as mp
    with mp.get_context("fork").Pool(PROCS) as pool:
        res = dict(pool.map(_background_one, list(DISTS)))
    dump("background.json", res)


# (dist, Kx, Kz, alpha, omega list)
def disp_cases():
    rng = np.random.RandomState(1234)

    def box(n, x0, x1, y0, y1):
        return list(rng.uniform(x0, x1, n) + 1j*rng.uniform(y0, y1, n))

    cases = []
    cases.append(("mrn_1", 100.0, 10.0, 0.0, [-2 - 2j, 2 - 2j, -2 + 2j, 2 + 2j]))   # test_psi_mode_mpi.py:108
    cases.append(("mrn_1e-2", 10.0, 1000.0, 0.0, box(10, -2, 2, 1e-8, 1) +
                  [-1.0062579942546936 + 1.3209787635623396e-05j, 0.3 - 0.2j, 0.0 + 0.5j]))
    cases.append(("mrn_1", 0.1, 10.0, 0.0, box(8, -2, 2, 1e-8, 1) + [-1.0048 + 1e-4j, -1.0048 + 1e-6j]))
    cases.append(("mrn_1", 10.0, 10.0, 0.0, box(6, -2, 2, 2e-7, 2) + [0.05 + 1e-3j, -0.3 + 1e-5j]))
    cases.append(("mrn_1e-1", 50.0, 100.0, 0.0, box(12, -7.5, 7.5, -12.5, 2.5) + [1.0 + 0j, -0.4 + 1e-7j]))
    cases.append(("mrn_1e-1", 80.0, 100.0, 4e-8, box(8, -2, 2, 1e-8, 1) +
                  [0.5682657887635131 + 0.021369419154787475j]))
    cases.append(("mrn_1", 30.0, 300.0, 1e-6, box(6, -2, 2, 1e-8, 1)))
    cases.append(("pl32_c10", 5.0, 50.0, 0.0, box(6, -2, 2, 1e-8, 1)))
    cases.append(("pl32_c10", 5.0, 50.0, 1e-5, box(4, -2, 2, 1e-8, 1)))
    cases.append(("bump", 200.0, 1000.0, 0.0, box(8, -2, 2, 1e-8, 1) +
                  [0.5061011996916127 + 0.6192442962323604j, 1.1 + 1e-4j]))
    cases.append(("bump", 12.0, 12.0, 1e-7, box(4, -2, 2, 1e-7, 1)))
    cases.append(("bumptail", 200.0, 1000.0, 0.0, box(8, -2, 2, 1e-8, 1)))
    cases.append(("bump_eps", 20.0, 100.0, 0.0, box(6, -2, 2, 1e-8, 1)))
    cases.append(("single", 30.0, 30.0, 0.0, box(6, -2, 2, 1e-8, 1) +
                  [0.34801869149969916 + 0.4190301951732903j]))
    return cases


def _disp_one(case):
    ns = refshim.load()
    name, kx, kz, alpha, ws = case
    pd = ns.psi_dispersion.PSIDispersion(**ref_init_kwargs(ns, DISTS[name]))
    D, M, poles = [], [], []
    for w in ws:
        D.append(complex(pd.calculate(w, kx, kz, alpha)))
        poles.append([float(np.real(p)) for p in pd.poles])
        m = pd.matrix_M(w)
        M.append([m[0, 0], m[0, 1], m[0, 2], m[1, 0], m[1, 1], m[1, 2], m[2, 2]])
    return dict(dist=name, kx=kx, kz=kz, alpha=alpha, w=cl(ws), D=cl(D),
                M=[cl(m) for m in M], poles=poles)


def sec_disp():
    import multiprocessing as mp
    with mp.get_context("fork").Pool(PROCS) as pool:
        res = pool.map(_disp_one, disp_cases(), chunksize=1)
    # the reference's own default (QUADPACK) path at its pinned points, for the record
    ns = refshim.load()
    pdq = ns.psi_dispersion.PSIDispersion(**ref_init_kwargs(ns, DISTS["mrn_1"], with_ts=False))
    ws = [-2 - 2j, 2 - 2j, -2 + 2j, 2 + 2j]
    quad = cl(pdq.calculate(np.asarray(ws), 100.0, 10.0, 0.0))
    dump("disp.json", dict(cases=res, quadpack_mrn_1_K100_10=dict(w=cl(ws), D=quad)))


# ------------------------------------------------------------------------------------------
def count_arglist(ns=None, ts=None):
    """test_complex_roots_mpi.py:32-70."""
    z_list = [-2.0 + 2e-7j, 2.0 + 2e-7j, 2.0 + 2.0j, -2.0 + 2.0j]
    ax = np.logspace(-1, 3, 3)
    kxg, kzg = np.meshgrid(ax, ax)
    items = []
    for kx, kz in zip(kxg.flatten(), kzg.flatten()):
        init = {'stokes_range': [1e-8, 1.0], 'dust_to_gas_ratio': 3.0,
                'size_distribution_power': 3.5}
        if ts is not None:
            init['tanhsinh_integrator'] = ts
        items.append({'PSIDispersion.__init__': init,
                      'dispersion': {'wave_number_x': float(kx), 'wave_number_z': float(kz),
                                     'viscous_alpha': 0.0},
                      'z_list': z_list, 'count_roots': {'tol': 0.1}})
    return items


def _count_one(item):
    ns = refshim.load()
    item = dict(item)
    init = dict(item['PSIDispersion.__init__'])
    init['tanhsinh_integrator'] = ns.tanhsinh.TanhSinhNoDeepCopy()
    disp = ns.psi_dispersion.PSIDispersion(**init).calculate
    f = lambda z: disp(z, **item['dispersion'])
    cp = ns.complex_roots.ClosedPath(f, item['z_list'])
    trace = []
    orig = cp.refine_select

    def traced(**kw):
        n = orig(**kw)
        trace.append(int(n))
        return n
    cp.refine_select = traced
    try:
        n = cp.count_roots(**item['count_roots'])
    except RuntimeError:
        n = -666
    return dict(kx=item['dispersion']['wave_number_x'], kz=item['dispersion']['wave_number_z'],
                count=int(n), n_points=int(len(cp.points)), batches=trace,
                points=cl(cp.points), f_points=cl(cp.f_points))


def sec_contour():
    import multiprocessing as mp
    items = count_arglist()
    t0 = time.time()
    with mp.get_context("fork").Pool(PROCS) as pool:
        res = pool.map(_count_one, items, chunksize=1)
    print("contour counts", [r['count'] for r in res], "in %.0f s" % (time.time() - t0))
    dump("contour.json", dict(z_list=cl(items[0]['z_list']), tol=0.1, cases=res))


# ------------------------------------------------------------------------------------------
def mode_cases():
    """test_psi_mode.py:32-189 (all six, with TanhSinh() everywhere so the device path is
    comparable; the QUADPACK originals are recorded as pins) + extra seeds."""
    base = dict(real_range=[-2, 2], imag_range=[1.0e-8, 1])
    cases = [
        dict(name="t0_single", dist="single", init=dict(base), K=(30, 30), alpha=0, seed=0,
             pin=[0.34801869149969916, 0.4190301951732903]),
        dict(name="t1_mrn_1e-2", dist="mrn_1e-2", init=dict(base), K=(10, 1000), alpha=0, seed=0,
             pin=[-1.0062579942546936, 1.3209787635623396e-05]),
        dict(name="t2_mrn_1", dist="mrn_1", init=dict(base), K=(0.1, 10), alpha=0, seed=0,
             pin=[-1.004879704650626, 0.0008946325671658642]),
        dict(name="t3_bump", dist="bump", init=dict(base, n_sample=15, max_zoom_domains=1),
             K=(200, 1000), alpha=0, seed=0, pin=[0.5061011996916127, 0.6192442962323604]),
        dict(name="t4_bumptail", dist="bumptail", init=dict(base, n_sample=15, max_zoom_domains=1),
             K=(200, 1000), alpha=0, seed=0, pin=[0.5067259648099907, 0.6200307252463185]),
        dict(name="t5_visc", dist="mrn_1e-1", init=dict(base, n_sample=15, max_zoom_domains=1),
             K=(80.0, 100.0), alpha=4e-8, seed=0, pin=[0.5682657887635131, 0.021369419154787475]),
    ]
    return cases


def _mode_one(case):
    ns = refshim.load()
    kw = ref_init_kwargs(ns, DISTS[case["dist"]])
    kw.update(case["init"])
    np.random.seed(case["seed"])
    pm = ns.psi_mode.PSIMode(**kw)
    roots = pm.calculate(wave_number_x=case["K"][0], wave_number_z=case["K"][1],
                         viscous_alpha=case["alpha"], guess_roots=case.get("guess", []))
    out = dict(case)
    out.update(roots=cl(roots), n_function_call=int(pm.n_function_call),
               z_sample=cl(pm.z_sample), f_sample=cl(pm.f_sample),
               mask=[int(v) for v in pm.ra.maskF], weights=cl(pm.ra.weights))
    return out


def sec_mode():
    import multiprocessing as mp
    with mp.get_context("fork").Pool(PROCS) as pool:
        res = pool.map(_mode_one, mode_cases(), chunksize=1)
    for r in res:
        print(r["name"], r["roots"], r["n_function_call"])
    dump("mode.json", dict(cases=res))


def mode_grid_arglist(n_axis, stokes_range, seed=2, n_sample=20, max_zoom=1):
    """Arg list of test_psi_mode_mpi.py:38-71 on an n_axis x n_axis logspace(-1,3) grid."""
    ax = np.logspace(-1, 3, n_axis)
    kxg, kzg = np.meshgrid(ax, ax)
    items = []
    for kx, kz in zip(kxg.flatten(), kzg.flatten()):
        items.append({'__init__': {'stokes_range': list(stokes_range), 'dust_to_gas_ratio': 3.0,
                                   'size_distribution_power': 3.5, 'real_range': [-2.0, 2.0],
                                   'imag_range': [1e-8, 1.0], 'n_sample': n_sample,
                                   'max_zoom_domains': max_zoom, 'tol': 1.0e-13,
                                   'clean_tol': 1e-4},
                      'calculate': {'wave_number_x': float(kx), 'wave_number_z': float(kz)},
                      'random_seed': seed})
    return items


def _grid_item(item):
    ns = refshim.load()
    import copy
    item = copy.deepcopy(item)
    item['__init__']['tanhsinh_integrator'] = ns.tanhsinh.TanhSinhNoDeepCopy()
    s = ns.psi_mode_mpi.MpiScheduler.__new__(ns.psi_mode_mpi.MpiScheduler)
    s.verbose = False
    s.rank = 1
    s.PSIModeArgs = {}
    t0 = time.time()
    roots = s.runcompute(item)
    return dict(kx=item['calculate']['wave_number_x'], kz=item['calculate']['wave_number_z'],
                roots=cl(roots), n_function_call=int(s.pm.n_function_call),
                seconds=time.time() - t0)


def sec_modegrid():
    """3x3 grid of test_psi_mode_mpi.py:38-104 and a 6x6 grid of the same family (config 3's
    distribution; 6 points per axis out of logspace(-1,3)), seed 2."""
    import multiprocessing as mp
    out = {}
    for n_axis in (3, 6):
        items = mode_grid_arglist(n_axis, [1e-8, 1.0])
        t0 = time.time()
        with mp.get_context("fork").Pool(PROCS) as pool:
            res = pool.map(_grid_item, items, chunksize=1)
        print("mode grid", n_axis, "in %.0f s" % (time.time() - t0),
              [len(r['roots']) for r in res], flush=True)
        out["grid%d" % n_axis] = dict(n_axis=n_axis, stokes_range=[1e-8, 1.0], seed=2, cases=res)
    dump("modegrid.json", out)


# ------------------------------------------------------------------------------------------
def sec_refine():
    """test_psi_grid_refine.py:35-77 (tiny config), scheduler replaced by a process pool."""
    ns = refshim.load()
    pgr = ns.psi_grid_refine
    pb = ns.power_bump.PowerBump(amin=1e-8, aP=0.1, aL=2.0/3.0*0.1, aR=0.156, bumpfac=2.0)
    sd = ns.sizedensity.SizeDensity(pb.sigma0, [1e-8, 0.156])
    sd.poles = [pb.get_discontinuity()]
    baseargs = {'__init__': {'stokes_range': [1e-8, 0.156], 'dust_to_gas_ratio': 10.0,
                             'size_density': sd, 'real_range': [-2.0, 2.0],
                             'imag_range': [1e-7, 1.0], 'n_sample': 20, 'max_zoom_domains': 1,
                             'tol': 1.0e-13, 'clean_tol': 1e-4,
                             'tanhsinh_integrator': ns.tanhsinh.TanhSinhNoDeepCopy()},
                'calculate': {'wave_number_x': None, 'wave_number_z': None, 'guess_roots': []},
                'random_seed': 2}
    ref = pgr.PSIGridRefiner.__new__(pgr.PSIGridRefiner)
    # constructor body without the MPI scheduler (psi_grid_refine.py:101-149)
    ref.baseargs = baseargs
    ref.nbase = (2, 2)
    ref.krange = (1.0, 1.1)
    ref.reruns = 0
    ref.batchname = 'golden'
    ref.verbose = False
    ref.wall_start = time.time()
    ref.comm = sys.modules['mpi4py'].MPI.COMM_WORLD
    ref.rank = 0
    ref.root = True
    ref.ms = refshim.SerialScheduler("mode", PROCS)
    ref.grids = []
    log = []
    orig_run = ref.ms.run

    def logged(arglist):
        r = orig_run(arglist)
        log.append([dict(kx=float(a['calculate']['wave_number_x']),
                         kz=float(a['calculate']['wave_number_z']),
                         seed=int(a['random_seed']),
                         max_zoom=int(a['__init__']['max_zoom_domains']),
                         guess=cl(a['calculate']['guess_roots']),
                         roots=cl(r[i])) for i, a in enumerate(arglist)])
        return r
    ref.ms.run = logged
    t0 = time.time()
    ref.run_basegrid()
    ref.fill_in_grid()
    g = ref.grids[-1]
    out = dict(seconds=time.time() - t0, batches=log,
               final=dict(Kx=g['Kx'].tolist(), Kz=g['Kz'].tolist(), runs=g['runs'].tolist(),
                          nguess=g['nguess'].tolist(),
                          results={str(k): cl(v) for k, v in g['results'].items()}))
    dump("refine.json", out)


def sec_modegrid16():
    """BASELINE config 3 parity sample: the 16x16 strided subset of the 256x256
    logspace(-1,3) (Kx,Kz) map, MRN mu=3, tau in [1e-8,1], n_sample 20, max_zoom 1, seed 2."""
    import multiprocessing as mp
    ax = np.logspace(-1, 3, 256)[8::16]
    kxg, kzg = np.meshgrid(ax, ax)
    items = []
    for kx, kz in zip(kxg.flatten(), kzg.flatten()):
        items.append({'__init__': {'stokes_range': [1e-8, 1.0], 'dust_to_gas_ratio': 3.0,
                                   'size_distribution_power': 3.5, 'real_range': [-2.0, 2.0],
                                   'imag_range': [1e-8, 1.0], 'n_sample': 20, 'max_zoom_domains': 1,
                                   'tol': 1.0e-13, 'clean_tol': 1e-4},
                      'calculate': {'wave_number_x': float(kx), 'wave_number_z': float(kz)},
                      'random_seed': 2})
    t0 = time.time()
    with mp.get_context("fork").Pool(PROCS) as pool:
        res = pool.map(_grid_item, items, chunksize=1)
    print("mode grid 16x16 strided in %.0f s" % (time.time() - t0), sum(len(r['roots']) > 0 for r in res),
          "pixels with roots", flush=True)
    dump("modegrid16.json", dict(axis=list(map(float, ax)), stokes_range=[1e-8, 1.0], seed=2, cases=res))


def sec_multistart():
    """PSIMode searches whose rational approximation has two or three zeros of VERY different Im inside the
    domain (MRN mu=3, tau in [1e-8, 0.1], seed 2): find_dispersion_roots polishes them one scalar at a time
    (psi_mode.py:340-352), so each secant run uses the tolerance of its own start, min(1.48e-8, 1e-5*|Im|) --
    n_function_call pins that (an implementation that takes the minimum over all starts iterates longer)."""
    import multiprocessing as mp
    ks = [(0.1, 16.68100537200059), (0.1, 1000.0), (5.994842503189409, 129.15496650148827),
          (5.994842503189409, 1000.0), (0.774263682681127, 5.994842503189409), (0.774263682681127, 2.1544346900318834)]
    items = [{'__init__': {'stokes_range': [1e-8, 0.1], 'dust_to_gas_ratio': 3.0, 'size_distribution_power': 3.5,
                           'real_range': [-2.0, 2.0], 'imag_range': [1e-8, 1.0], 'n_sample': 20,
                           'max_zoom_domains': 1, 'tol': 1.0e-13, 'clean_tol': 1e-4},
              'calculate': {'wave_number_x': kx, 'wave_number_z': kz}, 'random_seed': 2} for kx, kz in ks]
    with mp.get_context("fork").Pool(PROCS) as pool:
        res = pool.map(_grid_item, items, chunksize=1)
    for r in res:
        print(r['kx'], r['kz'], r['roots'], r['n_function_call'])
    dump("multistart.json", dict(stokes_range=[1e-8, 0.1], seed=2, cases=res))


def table_sigma(a):
    """A user-style size density no analytic family covers: MRN slope with a log-normal enhancement around a = 0.02
    and a steeper slope above a = 0.05 (continuous, kink at 0.05 -> SizeDensity.poles = [0.05])."""
    a = np.asarray(a, dtype=float)
    base = a**-0.5*(1.0 + 3.0*np.exp(-np.log(a/0.02)**2/0.5))
    return np.where(a > 0.05, base*(a/0.05)**-1.0, base)


def sec_table():
    """Opaque callable size density (sizedensity.py:25-48 with a user lambda): background, D(omega) at 10 points
    (inviscid and viscous) and two PSIMode searches, for the tabulated-size-density path of the product."""
    ns = refshim.load()
    sd = ns.sizedensity.SizeDensity(table_sigma, [1e-6, 0.2])
    sd.poles = [0.05]
    ts = ns.tanhsinh.TanhSinhNoDeepCopy()
    pd = ns.psi_dispersion.PSIDispersion(2.0, [1e-6, 0.2], size_density=sd, tanhsinh_integrator=ts)
    rng = np.random.RandomState(77)
    w = list(rng.uniform(-2, 2, 8) + 1j*rng.uniform(-1, 1, 8)) + [0.35 + 1e-4j, -0.8 + 2e-3j]
    out = dict(mu=2.0, stokes_range=[1e-6, 0.2], poles=[0.05], rhod=float(sd.rhod), vgx=float(pd.vgx),
               vgy=float(pd.vgy), cases=[])
    for kx, kz, alpha in ((8.0, 40.0, 0.0), (60.0, 300.0, 0.0), (8.0, 40.0, 1e-6)):
        D, M = [], []
        for x in w:
            D.append(complex(pd.calculate(x, kx, kz, alpha)))
            m = pd.matrix_M(x)
            M.append([m[0, 0], m[0, 1], m[0, 2], m[1, 0], m[1, 1], m[1, 2], m[2, 2]])
        out["cases"].append(dict(kx=kx, kz=kz, alpha=alpha, w=cl(w), D=cl(D), M=[cl(m) for m in M]))
    modes = []
    for kx, kz, seed in ((1.0, 100.0, 0), (30.0, 10.0, 0), (100.0, 1000.0, 0), (3.0, 30.0, 1)):
        np.random.seed(seed)
        pm = ns.psi_mode.PSIMode(2.0, [1e-6, 0.2], [-2, 2], [1e-8, 1], size_density=sd, tanhsinh_integrator=ts)
        roots = pm.calculate(wave_number_x=kx, wave_number_z=kz)
        modes.append(dict(kx=kx, kz=kz, seed=seed, roots=cl(roots), n_function_call=int(pm.n_function_call)))
        print(kx, kz, roots, pm.n_function_call, flush=True)
    out["modes"] = modes
    dump("table.json", out)


def sec_boundary():
    """Public methods of the boundary classes beyond calculate/count_roots: PSIDispersion.matrix_P
    and the stopping-time matrices, average/int_j, PSIMode.find_root (scalar and array) and add_extra_domain,
    ClosedPath.refine / refine_select / interweave -- values from the unmodified reference."""
    ns = refshim.load()
    ts = ns.tanhsinh.TanhSinhNoDeepCopy()
    pd = ns.psi_dispersion.PSIDispersion(3.0, [1e-8, 1.0], tanhsinh_integrator=ts)
    w, kx, kz, alpha = 0.3 + 0.2j, 10.0, 100.0, 1e-6
    out = dict(w=cl([w])[0], kx=kx, kz=kz, alpha=alpha)
    out["D"] = cl([complex(pd.calculate(w, kx, kz, alpha))])[0]
    out["P"] = [cl(row) for row in pd.matrix_P(w)]
    taus = [1e-6, 3e-3, 0.2, 0.9]
    out["taus"] = taus
    A = pd.matrix_inverse_A(w)
    AD = pd.matrix_AD(A, w)
    VAD = pd.matrix_VAD(AD, w)
    W = pd.matrix_W(w)
    ev = lambda mat: [[cl([complex(np.asarray(mat[i][j](np.asarray(taus)) + 0*np.asarray(taus))[k]) for k in range(4)])
                       for j in range(3)] for i in range(3)]
    out["Ainv"], out["AD"], out["VAD"], out["W"] = ev(A), ev(AD), ev(VAD), ev(W)
    out["int_j"] = [float(np.real(pd.int_j(0))), float(np.real(pd.int_j(1)))]
    out["average_tau2"] = cl([complex(pd.average(lambda x: x**2/(1 + 1j*x)))])[0]
    # PSIMode: calculate, then find_root on a scalar and on an array, then add_extra_domain
    np.random.seed(3)
    pm = ns.psi_mode.PSIMode(3.0, [1e-8, 1.0], [-2, 2], [1e-8, 1], tanhsinh_integrator=ts)
    roots = pm.calculate(wave_number_x=0.1, wave_number_z=10.0)
    out["mode_roots"] = cl(roots)
    n0 = pm.n_function_call
    r1 = pm.find_root(roots[0]*(1 + 1e-3), max_iter=50)
    n1 = pm.n_function_call
    starts = np.array([roots[0]*(1 - 2e-3), 0.5 + 0.3j, roots[0] + 1e-4])
    r2 = pm.find_root(starts.copy(), max_iter=30)
    n2 = pm.n_function_call
    r3 = pm.find_root(np.array([1.9 + 0.9j]), max_iter=8, select_inside_domain=False)
    out["find_root"] = dict(scalar_start=cl([roots[0]*(1 + 1e-3)])[0], scalar=cl([complex(r1)])[0], calls_scalar=n1 - n0,
                            array_start=cl(starts), array=cl(r2), calls_array=n2 - n1,
                            outside_start=cl([1.9 + 0.9j]), outside=cl(r3), calls_outside=pm.n_function_call - n2)
    n3, m3 = pm.n_function_call, len(pm.z_sample)
    pm.add_extra_domain(extra_domain_size=[0.05, 0.02], centre=roots[0])
    out["extra_domain"] = dict(z_new=cl(pm.z_sample[m3:]), f_new=cl(pm.f_sample[m3:]), calls=pm.n_function_call - n3)
    # ClosedPath
    f = lambda z: pd.calculate(z, 0.1, 10.0)
    z_list = [-2.0 + 2e-7j, 2.0 + 2e-7j, 2.0 + 2.0j, -2.0 + 2.0j]
    cp = ns.complex_roots.ClosedPath(f, z_list)
    added = [cp.refine_select(tol=0.1, max_step=0.1) for _ in range(3)]
    out["closed_path"] = dict(z_list=cl(z_list), added=added, points=cl(cp.points), f_points=cl(cp.f_points))
    cp2 = ns.complex_roots.ClosedPath(f, z_list)
    cp2.refine()
    out["closed_path"]["refined_points"] = cl(cp2.points)
    out["closed_path"]["interweave"] = [float(v) for v in cp2.interweave(np.array([1.0, 2.0, 3.0]), np.array([10.0, 20.0, 30.0]))]
    dump("boundary.json", out)


def _follow_func(z, k):
    """Toy relation for RootFollower: a cubic whose roots move with k and cross regions where ten secant
    iterations from the fixed start values are not enough (rejected steps)."""
    return z**3 - (1.0 + 0.5j)*z - (k + 0.2j)


def sec_guessgen():
    """RootFollower (complex_roots.py:428-515) on a toy relation and PSIModeTV /
    TerminalVelocitySolver (terminalvelocitysolver.py) -- values from the unmodified reference."""
    import contextlib
    import importlib
    import io
    ns = refshim.load()
    tv = importlib.import_module("psitools.terminalvelocitysolver")
    out = {}
    cases = []
    for k_start, k_want in ((0.0, [0.05, 0.2, 0.5, 1.0, 3.0]), (0.0, [-0.4, -0.1, 0.3, 2.0]), (1.0, [1.5, 40.0]),
                            (0.5, [0.5001, 0.6])):
        roots0 = np.roots([1.0, 0.0, -(1.0 + 0.5j), -(k_start + 0.2j)])
        z_start = complex(roots0[np.argmax(roots0.imag)])
        with contextlib.redirect_stdout(io.StringIO()):
            got = ns.complex_roots.RootFollower(_follow_func, z_start, k_start).calculate(np.array(k_want))
        cases.append(dict(k_start=k_start, k_want=k_want, z_start=cj(z_start), roots=cl(got)))
    out["follow"] = cases
    tvs = []
    for mu, tmax, beta, kx, kz, alpha in ((3.0, 0.01, -3.5, 46.0, 1.0, 0.0), (3.0, 0.1, -3.5, 10.0, 100.0, 0.0),
                                          (1.0, 0.1, -3.5, 30.0, 30.0, 1e-7), (10.0, 0.01, -3.0, 100.0, 20.0, 0.0),
                                          (0.5, 0.03, -4.5, 20.0, 5.0, 0.0), (3.0, 0.01, -5.5, 46.0, 1.0, 1e-8),
                                          (3.0, 1.0, -3.5, 1.0, 1000.0, 0.0)):
        tau_min = np.logspace(-8, np.log10(tmax) - 0.01, 7)
        with contextlib.redirect_stdout(io.StringIO()):
            w = tv.PSIModeTV(mu, tmax, beta).find_roots(tau_min, kx, kz, viscous_alpha=alpha)
            w1 = tv.PSIModeTV(mu, tmax, beta).find_roots(float(tau_min[0]), kx, kz, viscous_alpha=alpha)
        tvs.append(dict(mu=mu, tau_max=tmax, beta=beta, kx=kx, kz=kz, alpha=alpha, tau_min=[float(t) for t in tau_min],
                        omega=cl(w), omega_scalar=cj(w1)))
    out["psimodetv"] = tvs
    tau_min = np.logspace(-8, -2, 100)
    with contextlib.redirect_stdout(io.StringIO()):
        w = tv.TerminalVelocitySolver(3, 46, 1, tau_min, 0.01, -3.5).find_roots()
    out["tvsolver_test_0"] = cl(w[::9])
    dump("guessgen.json", out)


SECTIONS = dict(guessgen=sec_guessgen, boundary=sec_boundary, table=sec_table, multistart=sec_multistart, modegrid16=sec_modegrid16, nodes=sec_nodes, background=sec_background, disp=sec_disp, contour=sec_contour,
                mode=sec_mode, modegrid=sec_modegrid, refine=sec_refine)

if __name__ == "__main__":
    if not refshim.available():
        sys.exit("reference tree not available")
    for name in (sys.argv[1:] or list(SECTIONS)):
        t0 = time.time()
        SECTIONS[name]()
        print("section", name, "done in %.1f s" % (time.time() - t0), flush=True)
