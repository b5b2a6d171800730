#!/usr/bin/env python
"""bench.py -- D(omega) evaluations per second on BASELINE config 2.

Workload ("config2"): the dispersion relation on a 1024x1024 complex-omega grid
(Re in [-7.5, 7.5], Im in [-12.5, 2.5], ranges of psitools_examples/example_PSIModeMPIDispersion.py)
at fixed (Kx, Kz) = (50, 100), MRN power law mu = 3, tau in [1e-8, 0.1], TanhSinh(15, 12).
One step = one pass of kernel K1 over the whole grid (1,048,576 evaluations per GPU).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

N > 1 is launched with torch.distributed.run (one rank per GPU); rank r evaluates its own 1024^2
shard of the N-times denser omega grid over the same rectangle (the columns r, r+N, ... of the N-times
denser Re axis: every rank sees the same Im rows, i.e. equal-cost shards), no data-path collective, weak
scaling.  Rank 0 prints ONE JSON line.  `--impl reference` times the CPU path on all host cores over a
bounded sample of the same workload: the UNMODIFIED reference staged into oracle/_ref by oracle/stage_ref.py
(kind "reference") when it is there, else the oracle port of its algorithm (kind "port").

The same line carries the secondary workloads: `modes` (config 3, 256^2 PSIMode map), `config4` (adaptive
refinement to the 1089^2 map) and `config5` (64 distributions x 128^2 searches) -- `--no-extra` skips the last two.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOPS_PER_NODE = 300.0          # nominal FP64 flops per joint node evaluation
GRID = 1024
K_BASE = (50.0, 100.0)
MU, STOKES = 3.0, (1e-8, 0.1)


def omega_grid(n=GRID, rank=0, world=1):
    """Rank r's shard of the omega grid that is `world` times denser along Re over the same rectangle: the base
    n x n grid shifted by r/world of a grid step in Re.  The cost of a point is set by its Im row (near-axis and
    strongly damped rows are the expensive ones), so every rank gets the same rows: equal-cost shards."""
    re = np.linspace(-7.5, 7.5, n)
    im = np.linspace(-12.5, 2.5, n)
    re = re + (rank/world)*(re[1] - re[0])
    return (re[None, :] + 1j*im[:, None]).ravel()


# ---------------------------------------------------------------------------------------------
# CPU arm: oracle port on the host cores
# ---------------------------------------------------------------------------------------------
def _one_blas_thread():
    """One BLAS thread per worker process: the pool already uses every core, and LAPACK on tiny
    matrices crawls when each of the processes spins up a full thread team."""
    try:
        from threadpoolctl import threadpool_limits
        return threadpool_limits(limits=1)
    except ImportError:
        import contextlib
        return contextlib.nullcontext()


def _cpu_worker(args):
    from oracle import psi_oracle as po
    w, kx, kz = args
    disp = _cpu_worker.cache.get("d")
    if disp is None:
        disp = po.Dispersion(po.DustDist.power_law(MU, list(STOKES)), po.NodeTable())
        _cpu_worker.cache["d"] = disp
    with _one_blas_thread():
        return disp.calculate(w, kx, kz)


_cpu_worker.cache = {}


def ref_staged():
    return os.path.isfile(os.path.join(ROOT, "oracle", "_ref", "psitools", "psi_dispersion.py"))


def _ref_ns():
    """The unmodified reference from oracle/_ref (staged by oracle/stage_ref.py) through the test suite's import
    shim (stubs for mpi4py / h5py, np.int)."""
    ns = _ref_ns.cache.get("ns")
    if ns is None:
        os.environ["PSI_REFERENCE_ROOT"] = os.path.join(ROOT, "oracle", "_ref")
        sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
        import warnings
        warnings.simplefilter("ignore")
        import refshim
        refshim.REF_ROOT = os.environ["PSI_REFERENCE_ROOT"]
        ns = _ref_ns.cache["ns"] = refshim.load()
    return ns


_ref_ns.cache = {}


def _ref_worker(args):
    w, kx, kz = args
    ns = _ref_ns()
    pd = _ref_worker.cache.get("d")
    if pd is None:
        pd = ns.psi_dispersion.PSIDispersion(MU, list(STOKES), tanhsinh_integrator=ns.tanhsinh.TanhSinh())
        _ref_worker.cache["d"] = pd
    with _one_blas_thread():
        return np.asarray(pd.calculate(w, kx, kz))


_ref_worker.cache = {}


def cpu_sample_rate(n_points, cores, seed=0, pool=None, worker=None):
    """evals/s of a CPU implementation (default: the oracle port) on `cores` processes over a seeded random
    subset of the grid."""
    import multiprocessing as mp
    worker = worker or _cpu_worker
    w = omega_grid()
    idx = np.random.RandomState(seed).choice(w.size, size=n_points, replace=False)
    chunks = [c for c in np.array_split(w[idx], cores*4) if len(c)]
    own = pool is None
    if own:
        pool = mp.get_context("fork").Pool(cores)
        pool.map(worker, [(w[:1], K_BASE[0], K_BASE[1])]*cores)      # warm the workers
    t0 = time.perf_counter()
    res = pool.map(worker, [(c, K_BASE[0], K_BASE[1]) for c in chunks], chunksize=1)
    dt = time.perf_counter() - t0
    if own:
        pool.close()
        pool.join()
    assert np.all(np.isfinite(np.concatenate(res)))
    return n_points/dt, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    ref = ref_staged()
    worker, kind = (_ref_worker, "reference") if ref else (_cpu_worker, "port")
    steps = min(args.steps, 8)                        # a bounded sample: the whole run ends within a few minutes
    per_step = (16 if ref else 256)*cores             # ~1.5 s of work per step on either implementation
    pool = mp.get_context("fork").Pool(cores)
    pool.map(worker, [(omega_grid()[:1], K_BASE[0], K_BASE[1])]*cores)
    for i in range(min(args.warmup, 2)):
        cpu_sample_rate(max(cores, per_step//8), cores, seed=100 + i, pool=pool, worker=worker)
    t_total, n_total = 0.0, 0
    for i in range(steps):
        _, dt = cpu_sample_rate(per_step, cores, seed=i, pool=pool, worker=worker)
        t_total += dt
        n_total += per_step
    pool.close()
    pool.join()
    value = n_total/t_total
    sample = "%d random points of the 1024^2 omega grid per step, %d steps (%s)" % (
        per_step, steps, "unmodified reference from oracle/_ref" if ref else "oracle port")
    line = {"impl": "reference", "metric": "D(omega) evals/s", "value": value, "unit": "evals/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 2),
            "ms_per_step": 1e3*t_total/steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.gpus),
            "cpu_baseline": {"value": value, "unit": "evals/s", "cores": cores, "kind": kind,
                             "sample": sample},
            "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ---------------------------------------------------------------------------------------------
# secondary workload: BASELINE config 3, (Kx,Kz) modes solved per second
# ---------------------------------------------------------------------------------------------
MODE_INIT = dict(stokes_range=[1e-8, 1.0], dust_to_gas_ratio=3.0, size_distribution_power=3.5,
                 real_range=[-2.0, 2.0], imag_range=[1e-8, 1.0], n_sample=20, max_zoom_domains=1,
                 tol=1.0e-13, clean_tol=1e-4)


def mode_grid(n_axis, rank=0, world=1):
    """Kx, Kz of the n_axis^2 log-spaced map of test_psi_mode_mpi.py:51-53; rank r owns the pixels
    i = r (mod world) of the flattened map."""
    ax = np.logspace(-1, 3, n_axis)
    kxg, kzg = np.meshgrid(ax, ax)
    return kxg.ravel()[rank::world], kzg.ravel()[rank::world]


def _cpu_mode_worker(k):
    from oracle import psi_oracle as po
    item = {'__init__': dict(MODE_INIT), 'calculate': {'wave_number_x': k[0], 'wave_number_z': k[1]},
            'random_seed': 2}
    tab = _cpu_mode_worker.cache.setdefault("t", po.NodeTable())
    with _one_blas_thread():
        return len(po.run_mode_item(item, tab))


_cpu_mode_worker.cache = {}


def cpu_mode_rate(n_axis, cores):
    """modes/s of the oracle port of PSIMode.calculate on `cores` processes over a strided subset of
    the map (6 pixels per core, ~8 s)."""
    import multiprocessing as mp
    kx, kz = mode_grid(n_axis)
    want = min(len(kx), 6*cores)
    idx = np.linspace(0, len(kx) - 1, want).astype(int)
    with mp.get_context("fork").Pool(cores) as pool:
        t0 = time.perf_counter()
        pool.map(_cpu_mode_worker, [(float(kx[i]), float(kz[i])) for i in idx], chunksize=1)
        dt = time.perf_counter() - t0
    return want/dt, dt, want


def workload_config(n_gpus):
    return {"workload": "config2: D(omega) on a 1024x1024 complex-omega grid at fixed (Kx,Kz)=(50,100), "
                        "MRN mu=3 tau in [1e-8,0.1], TanhSinh(15,12)",
            "points_per_gpu": GRID*GRID,
            "sharding": "rank r owns the 1024^2 sub-grid shifted by r/N of a grid step along Re (columns "
                        "interleaved over ranks, same Im rows everywhere: equal-cost shards)",
            "parallelism": "dp%d" % n_gpus,
            "l2": "256 MiB buffer written between timed steps (L2 flush)",
            "exit_rule": "tanh-sinh level loop: reference rule + quadratic-collapse predictor, extra split point at "
                         "tau* = -1/Im w for damped frequencies (default); the same steps under the literal "
                         "reference rule are in literal_exit_rule"}


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.fh = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=self.fh, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.fh.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in open(self.path):
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.path)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)),
                       reasons=sorted(reasons), samples=len(sm))
        return out


def run_config4(barrier, world):
    """BASELINE config 4: PSIGridRefiner, PowerBump(amin 1e-8, aP 0.1, aL 2/3 aP, aR 1.56 aP, bumpfac 2), mu = 10,
    nbase 16, reruns 3, six fill-ins -> 1089^2 map; every sweep's runs are interleaved over the ranks and exchanged
    as device records.  Timed between barriers (max over ranks is taken by the caller)."""
    from psitools_public_b200 import PowerBump, PSIGridRefiner, SizeDensity, TanhSinhNoDeepCopy
    pb = PowerBump(amin=1e-8, aP=0.1, aL=2.0/3.0*0.1, aR=0.156, bumpfac=2.0)
    sd = SizeDensity(pb.sigma0, [1e-8, 0.156])
    sd.poles = [pb.get_discontinuity()]
    baseargs = {'__init__': {'stokes_range': [1e-8, 0.156], 'dust_to_gas_ratio': 10.0, 'size_density': sd,
                             'real_range': [-2.0, 2.0], 'imag_range': [1e-7, 1.0], 'n_sample': 20,
                             'max_zoom_domains': 1, 'tol': 1.0e-13, 'clean_tol': 1e-4,
                             'tanhsinh_integrator': TanhSinhNoDeepCopy()},
                'calculate': {'wave_number_x': None, 'wave_number_z': None, 'guess_roots': []}, 'random_seed': 2}
    import contextlib
    import io
    gr = PSIGridRefiner('bench', baseargs=baseargs, nbase=(16, 16), reruns=3, krange=(-1, 3))
    sched = {}
    inner = gr.ms.run_arrays

    def timed(*a, **k):
        t = time.perf_counter()
        out = inner(*a, **k)
        sched["scheduler"] = sched.get("scheduler", 0.0) + time.perf_counter() - t
        st = gr.ms.last_stats
        for key in ("prep_seconds", "batch_seconds", "exchange_seconds", "wait_seconds", "lib_setup_seconds",
                    "lib_call_seconds"):
            sched[key] = sched.get(key, 0.0) + (st.get(key) or 0.0)
        for key, v in (st.get("stage_s") or {}).items():
            sched["dev_" + key] = sched.get("dev_" + key, 0.0) + v
        sched["exchange"] = st.get("exchange") or "host-staged all_gather"
        return out
    gr.ms.run_arrays = timed
    with contextlib.redirect_stdout(io.StringIO()):        # untimed: a two-fill-in campaign (69^2 map) takes the one-time
        warm = PSIGridRefiner('warm', baseargs=baseargs, nbase=(16, 16), reruns=3, krange=(-1, 3))   # costs out
        warm.run_basegrid()                                # (kernel modules, communicator, first allocations)
        warm.fill_in_grid()
        warm.fill_in_grid()
        warm._close_device_map()
    sched.clear()
    import gc
    from psitools_public_b200 import _device as _dev
    extra_t = {"gc_seconds": 0.0, "device_map_seconds": 0.0}
    mark = [0.0]

    def on_gc(phase, info):
        if phase == "start":
            mark[0] = time.perf_counter()
        else:
            extra_t["gc_seconds"] += time.perf_counter() - mark[0]
    saved = {}
    for name in ("__init__", "sweep", "update", "close"):       # time spent in the device mirror of the cell queue
        fn = getattr(_dev.RefineMap, name)
        saved[name] = fn

        def timed_fn(self, *a, _fn=fn, _name=name.strip('_'), **k):
            t = time.perf_counter()
            try:
                return _fn(self, *a, **k)
            finally:
                extra_t["device_map_seconds"] += time.perf_counter() - t
                extra_t["map_" + _name] = extra_t.get("map_" + _name, 0.0) + time.perf_counter() - t
        setattr(_dev.RefineMap, name, timed_fn)
    gc.callbacks.append(on_gc)
    prof = None
    if os.environ.get("PSI_BENCH_PROFILE"):                # diagnostic: where the host time of the campaign goes
        import cProfile
        prof = cProfile.Profile()
    barrier()
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        if prof:
            prof.enable()
        gr.run_basegrid()
        for _ in range(6):
            gr.fill_in_grid()
        if prof:
            prof.disable()
    barrier()
    dt = time.perf_counter() - t0
    gc.callbacks.remove(on_gc)
    for name, fn in saved.items():
        setattr(_dev.RefineMap, name, fn)
    sched.update(extra_t)
    if prof and int(os.environ.get("RANK", "0")) == 0:
        import pstats
        pstats.Stats(prof, stream=sys.stderr).sort_stats('tottime').print_stats(14)
    g = gr.grids[-1]
    runs = sum(n for _, n, _ in gr.sweep_log) + gr.grids[0]['runs'].size
    return {"workload": "config4: PSIGridRefiner PowerBump mu=10, nbase 16, reruns 3, 6 fill-ins -> %dx%d map"
                        % g['runs'].shape, "seconds": dt, "psimode_runs": int(runs), "sweeps": len(gr.sweep_log),
            "pixels_with_roots": int(np.count_nonzero(gr.grid_arrays(g)['error'] > 0)),
            "rank0_seconds_in_scheduler": sched.get("scheduler"),
            "rank0_split": {k: (round(v, 3) if isinstance(v, float) else v) for k, v in sched.items()}}


def run_config5(barrier, world):
    """BASELINE config 5: 8 x 8 (mu, tau_max) MRN distributions x a 128^2 (Kx,Kz) map of PSIMode searches, ONE
    batch of 1,048,576 searches through MpiScheduler.run (TanhSinh(15,12))."""
    from psitools_public_b200 import TanhSinhNoDeepCopy, psi_mode_mpi
    ts = TanhSinhNoDeepCopy()
    ax = np.logspace(-1, 3, 128)
    kxg, kzg = np.meshgrid(ax, ax)
    items = []
    for mu in np.logspace(-1, 1, 8):
        for tmax in np.logspace(-3, 0, 8):
            init = {'stokes_range': [1e-8, float(tmax)], 'dust_to_gas_ratio': float(mu), 'real_range': [-2.0, 2.0],
                    'imag_range': [1e-8, 1.0], 'n_sample': 20, 'max_zoom_domains': 1, 'tanhsinh_integrator': ts}
            items += [{'__init__': init, 'calculate': {'wave_number_x': float(a), 'wave_number_z': float(b)},
                       'random_seed': 2} for a, b in zip(kxg.ravel(), kzg.ravel())]
    ms = psi_mode_mpi.MpiScheduler(time.time(), 1e9, verbose=False)
    ms.run(items[:64*world])
    barrier()
    t0 = time.perf_counter()
    res = ms.run(items)
    barrier()
    dt = time.perf_counter() - t0
    return {"workload": "config5: 8 x 8 (mu, tau_max) MRN distributions x 128^2 (Kx,Kz), TanhSinh(15,12), one batch",
            "metric": "(Kx,Kz) modes solved/s", "unit": "modes/s", "searches": len(items), "seconds": dt,
            "searches_with_roots": int(sum(len(v) > 0 for v in res.values())) if res is not None else None,
            "exchange": ms.last_stats.get("exchange"), "rank0_D_evals": ms.last_stats.get("evals")}


def adjudicate(pd, w_all, D_default, D_literal, picks):
    """Exact (40-digit, tests/truth_mp.py) D at the given grid points: how far the default path and the reference's
    literal rule are from it.  Skipped (None) when mpmath is not importable."""
    try:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import mpmath as mp
        import truth_mp
    except Exception:
        return None
    out = []
    t = truth_mp.TruthDispersion(pd.mu, [pd.taumin, pd.taumax], c=pd.c, j01=(pd.j0, pd.j1), rhod=pd.desc.rhod)
    for i in picks:
        w = complex(w_all[i])
        t.sd_poles = [mp.mpf(-1.0/w.imag)] if w.imag < 0 else []
        exact = t.calculate(w, K_BASE[0], K_BASE[1])
        out.append({"w": [w.real, w.imag], "default_rel_err": abs(D_default[i] - exact)/abs(exact),
                    "literal_rule_rel_err": abs(D_literal[i] - exact)/abs(exact)})
    return out


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    import __graft_entry__ as ge
    if local == 0:
        ge.build()                      # no-op when the in-tree libpsi_b200.so is up to date
    if world > 1:
        dist.barrier()
    from psitools_public_b200 import PSIDispersion, TanhSinh

    pd = PSIDispersion(MU, list(STOKES), tanhsinh_integrator=TanhSinh(), device=local)
    ctx = pd.ctx
    kx, kz = K_BASE
    w_host = torch.from_numpy(omega_grid(GRID, rank, world).view(np.float64).copy()).pin_memory()   # pinned
    n = GRID*GRID
    # a dedicated (non-default) stream: the C ABI treats a NULL stream as "the context's own"
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)

    # resident inputs/outputs in HBM
    w_dev = w_host.to(dev, non_blocking=True)
    D_dev = torch.empty(2*n, dtype=torch.float64, device=dev)
    par = torch.tensor([kx, kz, 0.0], dtype=torch.float64, device=dev)
    dist_dev = torch.tensor([pd.dist_id], dtype=torch.int32, device=dev)
    nodes_dev = torch.zeros(1, dtype=torch.int64, device=dev)
    flush = torch.empty(256*1024*1024, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()

    cmask = 1 << ctx.point_class(pd.dist_id, 0.0)

    def step_resident():
        flush.fill_(1)
        ctx.disp_eval_dev(n, 1, dist_dev.data_ptr(), par.data_ptr(), par.data_ptr() + 8,
                          par.data_ptr() + 16, w_dev.data_ptr(), D_dev.data_ptr(), 0,
                          nodes_dev.data_ptr(), cmask, stream.cuda_stream)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    fp64_peak = ctx.measure_fp64_peak(5) if rank == 0 else None

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()                 # before the warm-up: nvidia-smi needs tens of ms to come up
    for _ in range(args.warmup):
        step_resident()
    barrier()
    nodes_dev.zero_()
    launches0 = ctx.launch_count
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
           for _ in range(args.steps)]
    barrier()
    ev[0].record(stream)
    for i in range(args.steps):
        flush.fill_(1)
        kev[i][0].record(stream)
        ctx.disp_eval_dev(n, 1, dist_dev.data_ptr(), par.data_ptr(), par.data_ptr() + 8,
                          par.data_ptr() + 16, w_dev.data_ptr(), D_dev.data_ptr(), 0,
                          nodes_dev.data_ptr(), cmask, stream.cuda_stream)
        kev[i][1].record(stream)
    ev[1].record(stream)
    barrier()
    launches = ctx.launch_count - launches0
    ms_total = ev[0].elapsed_time(ev[1])
    ms_kernel = float(np.mean([a.elapsed_time(b) for a, b in kev]))
    node_evals = int(nodes_dev.item())
    clocks = None
    if rank == 0:
        # K steps of ~6 ms are over before nvidia-smi (50 ms period) has sampled twice: keep the GPU under the
        # SAME load (the same step, untimed) a little longer so that the clocks line has samples under this load
        t_hold = time.perf_counter()
        while time.perf_counter() - t_hold < 0.35:
            step_resident()
            torch.cuda.synchronize()
        clocks = sampler.stop()
        clocks["window"] = ("warm-up + timed steps + the same step repeated untimed for 0.35 s "
                            "(nvidia-smi -lms 50)")
    D_resident = D_dev.cpu().numpy().view(np.complex128)

    # the same steps under the reference's LITERAL exit rule (tanhsinh.py:120-126), for the record: the
    # default rule stops the level loop once the increments have collapsed quadratically
    literal = None
    if rank == 0:
        ctx.exit_rule = "reference"
        step_resident()
        torch.cuda.synchronize()
        nodes_dev.zero_()
        lev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(3)]
        for a, b in lev:
            flush.fill_(1)
            a.record(stream)
            ctx.disp_eval_dev(n, 1, dist_dev.data_ptr(), par.data_ptr(), par.data_ptr() + 8,
                              par.data_ptr() + 16, w_dev.data_ptr(), D_dev.data_ptr(), 0,
                              nodes_dev.data_ptr(), cmask, stream.cuda_stream)
            b.record(stream)
        torch.cuda.synchronize()
        lit_ms = float(np.mean([a.elapsed_time(b) for a, b in lev]))
        lit_nodes = int(nodes_dev.item())/3
        D_lit = D_dev.cpu().numpy().view(np.complex128)
        dev_rel = np.abs(D_lit - D_resident)/(np.abs(D_lit) + np.median(np.abs(D_lit)))
        w_all = w_host.numpy().view(np.complex128)
        # The default path splits the integral at tau* = -1/Im w where Im w < -1/taumax (csrc/psi_quad.cuh); next to
        # the branch cuts of D in that corner the REFERENCE's 12 levels are not converged, so the two rules differ
        # there.  Outside the corner they must agree; inside, the worst points are adjudicated by 40-digit arithmetic.
        corner = w_all.imag < -1.0/STOKES[1]
        off = np.flatnonzero(dev_rel > 1e-10)
        literal = {"evals_per_s_per_gpu": n/(lit_ms*1e-3), "kernel_ms": lit_ms, "node_evals_per_launch": lit_nodes,
                   "fp64_frac": FLOPS_PER_NODE*lit_nodes/(lit_ms*1e-3)/1e12/fp64_peak,
                   "max_rel_difference_of_D_outside_damped_corner": float(dev_rel[~corner].max()),
                   "points_differing_by_more_than_1e-10": int(len(off)),
                   "of_which_outside_damped_corner": int(np.count_nonzero(~corner[off])),
                   "max_rel_difference_of_D": float(dev_rel.max())}
        literal["adjudicated_by_40_digit_arithmetic"] = adjudicate(pd, w_all, D_resident, D_lit,
                                                                   off[np.argsort(-dev_rel[off])][:3])
        ctx.exit_rule = "predicted"
    nodes_dev.zero_()

    # end to end through the public API: host buffers, H2D + D2H inside the timed region
    w_np = w_host.numpy().view(np.complex128)
    out_pinned = torch.empty(2*n, dtype=torch.float64).pin_memory()
    out_np = out_pinned.numpy().view(np.complex128)
    for _ in range(min(args.warmup, 2)):
        pd.ctx.disp_eval(pd.dist_id, kx, kz, 0.0, w_np, out=out_np)
    barrier()
    t0 = time.perf_counter()
    e_ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
    e_ev[0].record(stream)
    for _ in range(args.steps):
        got = pd.ctx.disp_eval(pd.dist_id, kx, kz, 0.0, w_np, out=out_np)    # synchronous on return
    e_ev[1].record(stream)
    barrier()
    e2e_s = time.perf_counter() - t0
    assert np.array_equal(got, D_resident), "resident and end-to-end results differ"
    assert np.all(np.isfinite(got))

    # ---- secondary: PSIMode growth-rate map (BASELINE config 3), native lock-step engine ----
    modes = None
    if args.modes_grid > 0:
        from psitools_public_b200 import psi_mode_mpi
        from psitools_public_b200.tanhsinh import TanhSinhNoDeepCopy
        ts2 = TanhSinhNoDeepCopy()
        init = dict(MODE_INIT, tanhsinh_integrator=ts2)

        def arglist(n_axis):
            ax = np.logspace(-1, 3, n_axis)
            kxg, kzg = np.meshgrid(ax, ax)
            return [{'__init__': init, 'calculate': {'wave_number_x': float(a), 'wave_number_z': float(b)},
                     'random_seed': 2} for a, b in zip(kxg.ravel(), kzg.ravel())]

        ms = psi_mode_mpi.MpiScheduler(time.time(), 1e9, verbose=False)
        items = arglist(args.modes_grid)
        ms.run(arglist(min(32, args.modes_grid)))                # warm-up: library binding, communicator,
        ms.run(items)                                             # ... and one untimed pass (device buffers grow
                                                                  # to the batch size on first use)
        barrier()
        l0 = ctx.launch_count
        t0 = time.perf_counter()
        res = ms.run(items)              # the call a user makes (reference: {index: roots} on rank 0, None elsewhere)
        barrier()
        m_s = time.perf_counter() - t0
        st = ms.last_stats
        res = res if res is not None else {}
        modes = {"metric": "(Kx,Kz) modes solved/s", "unit": "modes/s", "seconds": m_s,
                 "workload": "config3: PSIMode on a %dx%d logspace(-1,3) (Kx,Kz) map, MRN mu=3 tau in "
                             "[1e-8,1], n_sample 20, max_zoom 1, seed 2; pixels interleaved over ranks, "
                             "one ncclAllGather of packed result records out of the library's device buffer"
                             % (args.modes_grid, args.modes_grid),
                 "pixels": len(items), "pixels_with_roots": int(sum(len(v) > 0 for v in res.values())),
                 "rank0_passes": st.get("passes"), "rank0_D_evals": st.get("evals"),
                 "rank0_engine_seconds": st.get("seconds"), "rank0_launches": ctx.launch_count - l0,
                 "rank0_host_logic_seconds": st.get("host_logic_seconds"),
                 "rank0_device_seconds": st.get("device_seconds"),
                 "rank0_D_evals_in_K3": st.get("polish_evals"), "host_threads": ctx.default_engine_threads(),
                 "engine": st.get("engine"), "exchange": st.get("exchange"), "rank0_stage_seconds": st.get("stage_s"),
                 "rank0_host_split": st.get("host_split")}
        if args.modes_compare:
            # same map through the host-AAA engine (LAPACK on the host cores + K1/K3), for comparison
            ms2 = psi_mode_mpi.MpiScheduler(time.time(), 1e9, verbose=False, engine="native")
            ms2.run_all(arglist(min(32, args.modes_grid)))
            barrier()
            t0 = time.perf_counter()
            res2 = ms2.run_all(items)
            barrier()
            modes["native_engine_seconds"] = time.perf_counter() - t0
            modes["native_engine_pixels_with_roots"] = int(sum(len(v) > 0 for v in res2.values()))

    # ---- secondary: BASELINE config 4 (adaptive refinement to the 1089^2 map) and config 5 (parameter sweep) ----
    config4 = config5 = None
    if not args.no_extra:
        config4 = run_config4(barrier, world)
        config5 = run_config5(barrier, world)

    t = torch.tensor([ms_total, ms_kernel, e2e_s, modes["seconds"] if modes else 0.0,
                      config4["seconds"] if config4 else 0.0, config5["seconds"] if config5 else 0.0],
                     dtype=torch.float64, device=dev)
    cnt = torch.tensor([float(node_evals), float(launches)], dtype=torch.float64, device=dev)
    per_rank = torch.zeros(2*world, dtype=torch.float64, device=dev)
    per_rank[2*rank], per_rank[2*rank + 1] = ms_kernel, float(node_evals)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        dist.all_reduce(per_rank, op=dist.ReduceOp.SUM)
    ms_total, ms_kernel_max, e2e_s, modes_s, c4_s, c5_s = [float(x) for x in t.tolist()]
    node_evals_all, launches_all = [float(x) for x in cnt.tolist()]
    rank_ms = per_rank[0::2].tolist()
    rank_nodes = per_rank[1::2].tolist()

    if rank == 0:
        total_evals = float(n)*world*args.steps
        value = total_evals/(ms_total*1e-3)
        # per GPU: ALL ranks' node evaluations over the slowest rank's kernel time, divided by the number of GPUs
        tf = FLOPS_PER_NODE*(node_evals_all/args.steps)/(ms_kernel_max*1e-3)/1e12/world
        traffic = None
        try:                       # DRAM bytes per K1 launch from the committed ncu --set full capture
            with open(os.path.join(ROOT, "profiles", "k1_traffic.json")) as fh:
                traffic = json.load(fh)["traffic_bytes_per_launch"]
        except Exception:
            pass
        roof = {"bound": "fp64", "achieved": tf, "peak": fp64_peak, "unit": "TFLOP/s",
                "frac": tf/fp64_peak if fp64_peak else None, "traffic": traffic,
                "kernel": "disp_eval_thread_kernel (thread per point, levels 0-7) + disp_eval_cta_kernel / disp_eval_kernel (block / warp per point, "
                          "the points left over); timed together; per GPU", "kernel_ms": ms_kernel_max,
                "kernel_ms_per_rank": [round(x, 4) for x in rank_ms],
                "node_evals_per_launch": node_evals_all/args.steps/world,
                "node_evals_per_launch_per_rank": [x/args.steps for x in rank_nodes],
                "flops_per_node_eval": FLOPS_PER_NODE,
                "peak_source": "measured live by dfma_peak_kernel (MEASURED_PEAKS.json has no FP64 entry)",
                "hbm_gbs_algorithmic": 32.0*n/(ms_kernel_max*1e-3)/1e9}
        cores = os.cpu_count() or 1
        cpu = None
        if world == 1 and not args.no_cpu:
            n_cpu = min(n, 1500*cores)                  # ~10 s of CPU work on all cores
            rate, dt = cpu_sample_rate(n_cpu, cores)
            cpu = {"value": rate, "unit": "evals/s", "cores": cores, "kind": "port",
                   "sample": "%d random points of the 1024^2 omega grid (%.1f s)" % (n_cpu, dt)}
            if ref_staged():                            # the unmodified reference itself (oracle/_ref), ~10 s
                n_ref = 100*cores
                rrate, rdt = cpu_sample_rate(n_ref, cores, worker=_ref_worker)
                cpu = {"value": rrate, "unit": "evals/s", "cores": cores, "kind": "reference",
                       "sample": "%d random points of the 1024^2 omega grid (%.1f s), unmodified reference staged in "
                                 "oracle/_ref" % (n_ref, rdt),
                       "port": {"value": rate, "sample": "%d points (%.1f s)" % (n_cpu, dt)}}
        if modes:
            modes["seconds"] = modes_s
            modes["value"] = modes["pixels"]/modes_s
            if cpu is not None:
                rate, dt, npx = cpu_mode_rate(args.modes_grid, cores)
                modes["cpu_baseline"] = {"value": rate, "unit": "modes/s", "cores": cores, "kind": "port",
                                         "sample": "%d pixels strided over the map (%.1f s)" % (npx, dt)}
        line = {"metric": "D(omega) evals/s", "value": value, "unit": "evals/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total/args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": workload_config(world),
                "e2e": {"value": total_evals/e2e_s, "unit": "evals/s",
                        "h2d_bytes_per_step": 16*n + 28, "d2h_bytes_per_step": 16*n + 8},
                "gpu_launches": int(launches_all), "roofline": roof, "cpu_baseline": cpu,
                "clocks": clocks, "literal_exit_rule": literal, "modes": modes}
        if config4:
            config4["seconds"] = c4_s
            config4["runs_per_s"] = config4["psimode_runs"]/c4_s
            line["config4"] = config4
        if config5:
            config5["seconds"] = c5_s
            config5["value"] = config5["searches"]/c5_s
            line["config5"] = config5
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300, help="timed steps (default: a timed region of about 1 s)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extra", action="store_true", help="skip the config4 / config5 workloads")
    ap.add_argument("--modes-compare", action="store_true",
                    help="also time the PSIMode map with the host-AAA (native) engine")
    ap.add_argument("--modes-grid", type=int, default=256,
                    help="axis length of the secondary PSIMode map (0 = skip)")
    args = ap.parse_args()
    # stdout carries exactly ONE line, the JSON record: everything else a library prints there while the
    # benchmark runs (NCCL's version banner, for one) is sent to stderr
    sys.stdout.flush()
    global _RESULT_FD
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


_RESULT_FD = None


def emit(line):
    text = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(text.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, text)


if __name__ == "__main__":
    main()
