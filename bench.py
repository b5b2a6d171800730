#!/usr/bin/env python
"""bench.py -- D(omega) evaluations per second on BASELINE config 2.

Workload ("config2"): the dispersion relation on a 1024x1024 complex-omega grid
(Re in [-7.5, 7.5], Im in [-12.5, 2.5], ranges of psitools_examples/example_PSIModeMPIDispersion.py)
at fixed (Kx, Kz) = (50, 100), MRN power law mu = 3, tau in [1e-8, 0.1], TanhSinh(15, 12).
One step = one pass of kernel K1 over the whole grid (1,048,576 evaluations per GPU).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

N > 1 is launched with torch.distributed.run (one rank per GPU); rank r evaluates its own 1024^2
shard of the N-times denser omega grid over the same rectangle (the base grid shifted by r/N of a
grid step), no data-path collective, weak scaling.  Rank 0 prints ONE JSON line.  `--impl reference` times the CPU path (the oracle port of
the reference's algorithm -- the Python reference itself cannot travel to the GPU box) on all host
cores over a bounded sample of the same workload.
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
    """Rank r's shard of the (n*world)-times denser omega grid over the same rectangle: the base
    n x n grid shifted by r/world of a grid step in Re and Im (interleaved shards of equal cost)."""
    re = np.linspace(-7.5, 7.5, n)
    im = np.linspace(-12.5, 2.5, n)
    re = re + (rank/world)*(re[1] - re[0])
    im = im + (rank/world)*(im[1] - im[0])
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


def cpu_sample_rate(n_points, cores, seed=0, pool=None):
    """evals/s of the oracle port on `cores` processes over a seeded random subset of the grid."""
    import multiprocessing as mp
    w = omega_grid()
    idx = np.random.RandomState(seed).choice(w.size, size=n_points, replace=False)
    chunks = [c for c in np.array_split(w[idx], cores*4) if len(c)]
    own = pool is None
    if own:
        pool = mp.get_context("fork").Pool(cores)
        pool.map(_cpu_worker, [(w[:1], K_BASE[0], K_BASE[1])]*cores)      # warm the workers
    t0 = time.perf_counter()
    res = pool.map(_cpu_worker, [(c, K_BASE[0], K_BASE[1]) for c in chunks], chunksize=1)
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
    per_step = 256*cores
    pool = mp.get_context("fork").Pool(cores)
    pool.map(_cpu_worker, [(omega_grid()[:1], K_BASE[0], K_BASE[1])]*cores)
    for i in range(args.warmup):
        cpu_sample_rate(max(cores, per_step//8), cores, seed=100 + i, pool=pool)
    t_total, n_total = 0.0, 0
    for i in range(args.steps):
        _, dt = cpu_sample_rate(per_step, cores, seed=i, pool=pool)
        t_total += dt
        n_total += per_step
    pool.close()
    pool.join()
    value = n_total/t_total
    sample = "%d random points of the 1024^2 omega grid per step, %d steps" % (per_step, args.steps)
    line = {"impl": "reference", "metric": "D(omega) evals/s", "value": value, "unit": "evals/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3*t_total/args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.gpus),
            "cpu_baseline": {"value": value, "unit": "evals/s", "cores": cores, "kind": "port",
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
            "sharding": "rank r owns the 1024^2 sub-grid shifted by r/N of a grid step (interleaved shards)",
            "parallelism": "dp%d" % n_gpus,
            "l2": "256 MiB buffer written between timed steps (L2 flush)",
            "exit_rule": "tanh-sinh level loop: reference rule + quadratic-collapse predictor (default); the "
                         "same steps under the literal reference rule are in literal_exit_rule"}


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
        dev_rel = np.abs(D_lit - D_resident)/np.abs(D_lit)
        literal = {"evals_per_s_per_gpu": n/(lit_ms*1e-3), "kernel_ms": lit_ms, "node_evals_per_launch": lit_nodes,
                   "fp64_frac": FLOPS_PER_NODE*lit_nodes/(lit_ms*1e-3)/1e12/fp64_peak,
                   "max_rel_difference_of_D": float(dev_rel.max()),
                   "points_differing_by_more_than_1e-12": int(np.count_nonzero(dev_rel > 1e-12))}
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
        ms.run_all(arglist(min(32, args.modes_grid)))            # warm-up: library binding, communicator,
        ms.run_all(items)                                         # ... and one untimed pass (device buffers grow
                                                                  # to the batch size on first use)
        barrier()
        l0 = ctx.launch_count
        t0 = time.perf_counter()
        res = ms.run_all(items)                                   # the call a user makes; all ranks share the map
        barrier()
        m_s = time.perf_counter() - t0
        st = ms.last_stats
        modes = {"metric": "(Kx,Kz) modes solved/s", "unit": "modes/s", "seconds": m_s,
                 "workload": "config3: PSIMode on a %dx%d logspace(-1,3) (Kx,Kz) map, MRN mu=3 tau in "
                             "[1e-8,1], n_sample 20, max_zoom 1, seed 2; pixels interleaved over ranks, "
                             "one all-gather of the roots" % (args.modes_grid, args.modes_grid),
                 "pixels": len(items), "pixels_with_roots": int(sum(len(v) > 0 for v in res.values())),
                 "rank0_passes": st.get("passes"), "rank0_D_evals": st.get("evals"),
                 "rank0_engine_seconds": st.get("seconds"), "rank0_launches": ctx.launch_count - l0,
                 "rank0_host_logic_seconds": st.get("host_logic_seconds"),
                 "rank0_device_seconds": st.get("device_seconds"),
                 "rank0_D_evals_in_K3": st.get("polish_evals"), "host_threads": ctx.default_engine_threads(),
                 "engine": st.get("engine")}
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

    t = torch.tensor([ms_total, ms_kernel, e2e_s, modes["seconds"] if modes else 0.0],
                     dtype=torch.float64, device=dev)
    cnt = torch.tensor([float(node_evals), float(launches)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    ms_total, ms_kernel, e2e_s, modes_s = [float(x) for x in t.tolist()]
    node_evals_all, launches_all = [float(x) for x in cnt.tolist()]

    if rank == 0:
        total_evals = float(n)*world*args.steps
        value = total_evals/(ms_total*1e-3)
        tf = FLOPS_PER_NODE*(node_evals/args.steps)/(ms_kernel*1e-3)/1e12     # rank 0's kernel
        traffic = None
        try:                       # DRAM bytes per K1 launch from the committed ncu --set full capture
            with open(os.path.join(ROOT, "profiles", "k1_traffic.json")) as fh:
                traffic = json.load(fh)["traffic_bytes_per_launch"]
        except Exception:
            pass
        roof = {"bound": "fp64", "achieved": tf, "peak": fp64_peak, "unit": "TFLOP/s",
                "frac": tf/fp64_peak if fp64_peak else None, "traffic": traffic,
                "kernel": "disp_eval_thread_kernel (thread per point, levels 0-8) + disp_eval_kernel (warp per point, "
                          "the points left over); timed together", "kernel_ms": ms_kernel,
                "node_evals_per_launch": node_evals/args.steps, "flops_per_node_eval": FLOPS_PER_NODE,
                "peak_source": "measured live by dfma_peak_kernel (MEASURED_PEAKS.json has no FP64 entry)",
                "hbm_gbs_algorithmic": 32.0*n/(ms_kernel*1e-3)/1e9}
        cores = os.cpu_count() or 1
        cpu = None
        if world == 1 and not args.no_cpu:
            n_cpu = min(n, 2000*cores)                  # ~12 s of CPU work on all cores
            rate, dt = cpu_sample_rate(n_cpu, cores)
            cpu = {"value": rate, "unit": "evals/s", "cores": cores, "kind": "port",
                   "sample": "%d random points of the 1024^2 omega grid (%.1f s)" % (n_cpu, dt)}
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
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
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
