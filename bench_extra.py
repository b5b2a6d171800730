#!/usr/bin/env python
"""Secondary workloads (not part of the bench.py contract): root-count map (kernel K2) and adaptive
grid refinement (BASELINE config 4 family) on one GPU.  Prints one JSON line per workload."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

WORLD = int(os.environ.get("WORLD_SIZE", "1"))
RANK = int(os.environ.get("RANK", "0"))


def count_map(n_axis):
    """example_psicountgrid.py family: roots of D inside [-2,2]x[2e-7,2] on an n_axis^2 log K grid."""
    from psitools_public_b200 import TanhSinhNoDeepCopy, complex_roots_mpi
    ts = TanhSinhNoDeepCopy()
    z_list = [-2.0 + 2e-7j, 2.0 + 2e-7j, 2.0 + 2.0j, -2.0 + 2.0j]
    ax = np.logspace(-1, 3, n_axis)
    kxg, kzg = np.meshgrid(ax, ax)
    init = {'stokes_range': [1e-8, 1.0], 'dust_to_gas_ratio': 3.0, 'size_distribution_power': 3.5,
            'tanhsinh_integrator': ts}
    items = [{'PSIDispersion.__init__': init,
              'dispersion': {'wave_number_x': float(a), 'wave_number_z': float(b), 'viscous_alpha': 0.0},
              'z_list': z_list, 'count_roots': {'tol': 0.1}} for a, b in zip(kxg.ravel(), kzg.ravel())]
    ms = complex_roots_mpi.MpiScheduler(time.time(), 1e9, verbose=False)
    ms.run_all(items[:8])
    t0 = time.perf_counter()
    res = ms.run_all(items)
    dt = time.perf_counter() - t0
    counts = np.array([res[i] for i in range(len(items))])
    st = ms.last_stats
    print(json.dumps({"workload": "root-count map %dx%d (K2 contour_kernel)" % (n_axis, n_axis),
                      "seconds": dt, "contours_per_s": len(items)/dt, "D_evals": st.get("evals"),
                      "contour_points": st.get("points"), "failed(-666)": int(np.sum(counts == -666)),
                      "histogram": {str(k): int(v) for k, v in zip(*np.unique(counts, return_counts=True))}}),
          flush=True)


def refine(nbase, fills, reruns):
    """test_psi_grid_refine.py's PowerBump distribution on the default krange (-1, 3)."""
    from psitools_public_b200 import PowerBump, PSIGridRefiner, SizeDensity, TanhSinhNoDeepCopy
    pb = PowerBump(amin=1e-8, aP=0.1, aL=2.0/3.0*0.1, aR=0.156, bumpfac=2.0)
    sd = SizeDensity(pb.sigma0, [1e-8, 0.156])
    sd.poles = [pb.get_discontinuity()]
    baseargs = {'__init__': {'stokes_range': [1e-8, 0.156], 'dust_to_gas_ratio': 10.0, 'size_density': sd,
                             'real_range': [-2.0, 2.0], 'imag_range': [1e-7, 1.0], 'n_sample': 20,
                             'max_zoom_domains': 1, 'tol': 1.0e-13, 'clean_tol': 1e-4,
                             'tanhsinh_integrator': TanhSinhNoDeepCopy()},
                'calculate': {'wave_number_x': None, 'wave_number_z': None, 'guess_roots': []},
                'random_seed': 2}
    gr = PSIGridRefiner('bench', baseargs=baseargs, nbase=(nbase, nbase), reruns=reruns, krange=(-1, 3))
    sched = {"s": 0.0}
    per_sweep = []                                # [items, seconds in the scheduler, batch, exchange] per sweep
    inner = gr.ms.run_arrays                      # time spent inside the scheduler (device + exchange)

    def timed(*a, **k):
        t = time.perf_counter()
        out = inner(*a, **k)
        sched["s"] += time.perf_counter() - t
        st = gr.ms.last_stats
        per_sweep.append([len(a[1]), round(time.perf_counter() - t, 4), round(st.get("batch_seconds") or 0.0, 4),
                          round(st.get("exchange_seconds") or 0.0, 4),
                          {k: round(v, 4) for k, v in (st.get("stage_s") or {}).items()}, st.get("function_calls"), st.get("calls_pct")])
        for key in ("prep_seconds", "batch_seconds", "exchange_seconds", "wait_seconds", "lib_setup_seconds",
                    "lib_call_seconds"):
            sched[key] = sched.get(key, 0.0) + (st.get(key) or 0.0)
        for key, v in (st.get("stage_s") or {}).items():
            sched["dev_" + key] = sched.get("dev_" + key, 0.0) + v
        sched["redone"] = sched.get("redone", 0) + (st.get("redone_for_root_slots") or 0)
        sched["max_roots"] = max(sched.get("max_roots", 0), st.get("max_roots") or 0)
        return out
    gr.ms.run_arrays = timed
    t0 = time.perf_counter()
    gr.run_basegrid()
    times = [time.perf_counter() - t0]
    for _ in range(fills):
        gr.fill_in_grid()
        times.append(time.perf_counter() - t0)
    g = gr.grids[-1]
    runs = sum(n for _, n, _ in gr.sweep_log) + gr.grids[0]['runs'].size
    with_roots = int(np.count_nonzero(gr.grid_arrays(g)['error'] > 0))
    if RANK == 0:
        print(json.dumps({"workload": "PSIGridRefiner PowerBump nbase=%d, %d fill-ins, reruns=%d -> %dx%d map"
                                      % (nbase, fills, reruns, g['runs'].shape[0], g['runs'].shape[1]),
                          "n_gpus": WORLD, "seconds_cumulative": times, "psimode_runs": int(runs),
                          "runs_per_s": runs/times[-1], "sweeps": len(gr.sweep_log),
                          "seconds_in_scheduler": sched["s"], "seconds_host_bookkeeping": times[-1] - sched["s"],
                          "scheduler_split": {k: round(v, 3) for k, v in sched.items() if k != "s"},
                          "pixels_with_roots": with_roots,
                          **({"per_sweep": per_sweep} if os.environ.get("PSI_BENCH_PER_SWEEP") else {})}), flush=True)


def sweep(n_mu, n_tau, n_axis, level):
    """BASELINE config 5 family: n_mu x n_tau (mu, tau_max) MRN distributions x an n_axis^2 (Kx,Kz) map of
    PSIMode solves in ONE batch, tanh-sinh table (15, level)."""
    from psitools_public_b200 import TanhSinhNoDeepCopy, psi_mode_mpi
    ts = TanhSinhNoDeepCopy(15, level)
    ax = np.logspace(-1, 3, n_axis)
    kxg, kzg = np.meshgrid(ax, ax)
    items = []
    for mu in np.logspace(-1, 1, n_mu):
        for tmax in np.logspace(-3, 0, n_tau):
            init = {'stokes_range': [1e-8, float(tmax)], 'dust_to_gas_ratio': float(mu),
                    'real_range': [-2.0, 2.0], 'imag_range': [1e-8, 1.0], 'n_sample': 20,
                    'max_zoom_domains': 1, 'tanhsinh_integrator': ts}
            items += [{'__init__': init, 'calculate': {'wave_number_x': float(a), 'wave_number_z': float(b)},
                       'random_seed': 2} for a, b in zip(kxg.ravel(), kzg.ravel())]
    ms = psi_mode_mpi.MpiScheduler(time.time(), 1e9, verbose=False)
    ms.run_all(items[:64])
    t0 = time.perf_counter()
    res = ms.run_all(items)
    dt = time.perf_counter() - t0
    print(json.dumps({"workload": "parameter sweep: %d x %d (mu, tau_max) distributions x %dx%d (Kx,Kz), "
                                  "TanhSinh(15,%d), one batch" % (n_mu, n_tau, n_axis, n_axis, level),
                      "searches": len(items), "seconds": dt, "modes_per_s": len(items)/dt,
                      "D_evals": ms.last_stats.get("evals"), "engine": ms.last_stats.get("engine"),
                      "searches_with_roots": int(sum(len(v) > 0 for v in res.values()))}), flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--count-grid", type=int, default=32)
    ap.add_argument("--nbase", type=int, default=16)
    ap.add_argument("--fills", type=int, default=2)
    ap.add_argument("--reruns", type=int, default=3)
    ap.add_argument("--sweep", type=int, nargs=4, metavar=("N_MU", "N_TAU", "N_AXIS", "LEVEL"), default=None)
    a = ap.parse_args()
    if WORLD > 1:                                  # torchrun: one rank per GPU, NCCL for the exchanges
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl")
        warm = torch.zeros(8, device="cuda")
        dist.all_gather([torch.empty_like(warm) for _ in range(WORLD)], warm)    # communicator set-up is
        dist.all_reduce(warm)                                                    # not part of any workload
        torch.cuda.synchronize()
        os.dup2(2, 1) if RANK else None            # only rank 0 reports
    if a.sweep:
        sweep(*a.sweep)
    if a.count_grid > 0:
        count_map(a.count_grid)
    if a.nbase > 0:
        refine(a.nbase, a.fills, a.reruns)
    if WORLD > 1:
        dist.barrier()
        dist.destroy_process_group()
