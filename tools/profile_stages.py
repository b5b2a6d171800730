"""Where does the time of a PSIMode batch go?  Runs (a) the 256^2 map of BASELINE config 3 restricted to
the items one rank of an N-GPU job would own (i mod N == 0) for N = 1, 2, 4, 8 and (b) a PSIGridRefiner
run, printing wall time, the per-stage stream times reported by psi_mode_batch_device and the host-side
share of MpiScheduler.run_all."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from psitools_public_b200 import psi_mode_mpi, PSIGridRefiner, PowerBump, SizeDensity, TanhSinhNoDeepCopy  # noqa: E402

ts = TanhSinhNoDeepCopy()
INIT = {'stokes_range': [1e-8, 1.0], 'dust_to_gas_ratio': 3.0, 'real_range': [-2, 2],
        'imag_range': [1e-8, 1], 'n_sample': 20, 'max_zoom_domains': 1, 'tol': 1e-13, 'clean_tol': 1e-4,
        'tanhsinh_integrator': ts}


def arglist(n_axis):
    ax = np.logspace(-1, 3, n_axis)
    kxg, kzg = np.meshgrid(ax, ax)
    return [{'__init__': INIT, 'calculate': {'wave_number_x': float(a), 'wave_number_z': float(b)},
             'random_seed': 2} for a, b in zip(kxg.ravel(), kzg.ravel())]


def fmt(st):
    return " ".join("%s %.3f" % kv for kv in st.items())


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "map"
    ms = psi_mode_mpi.MpiScheduler(time.time(), 1e9, verbose=False)
    ms.run_all(arglist(16))
    if what in ("map", "all"):
        items = arglist(int(os.environ.get("PSI_MAP", "256")))
        for n in (1, 2, 4, 8):
            sub = items[::n]
            for rep in range(2):
                t0 = time.perf_counter()
                ms.run_all(sub)
                dt = time.perf_counter() - t0
            st = ms.last_stats
            print("rank share 1/%d: %d items, run_all %.3f s, engine %.3f s | stages: %s" %
                  (n, len(sub), dt, st['seconds'], fmt(st.get('stage_s') or {})), flush=True)
    if what in ("refine", "all"):
        fills = int(os.environ.get("PSI_FILLS", "3"))
        pb = PowerBump(amin=1e-8, aP=0.1, aL=2/3*0.1, aR=1.56*0.1, bumpfac=2.0)
        sd = SizeDensity(pb.sigma0, [1e-8, 0.156])
        sd.poles = [pb.get_discontinuity()]
        base = {'__init__': {'stokes_range': [1e-8, 0.156], 'dust_to_gas_ratio': 10.0, 'size_density': sd,
                             'real_range': [-2.0, 2.0], 'imag_range': [1e-7, 1.0], 'n_sample': 20,
                             'max_zoom_domains': 1, 'tol': 1e-13, 'clean_tol': 1e-4,
                             'tanhsinh_integrator': ts},
                'calculate': {'wave_number_x': None, 'wave_number_z': None, 'guess_roots': []},
                'random_seed': 2}

        class Sched(psi_mode_mpi.MpiScheduler):
            def run_all(self, al):
                t0 = time.perf_counter()
                res = super().run_all(al)
                st = self.last_stats
                ng = [len(a['calculate'].get('guess_roots', ())) for a in al]
                print("  batch %6d items (guesses max %d mean %.1f): %.3f s, engine %.3f | %s" %
                      (len(al), max(ng), float(np.mean(ng)), time.perf_counter() - t0, st['seconds'],
                       fmt(st.get('stage_s') or {})), flush=True)
                return res

        gr = PSIGridRefiner('/tmp/psi_prof', baseargs=base, nbase=(16, 16), reruns=3, verbose=False,
                            scheduler=Sched(time.time(), 1e9, verbose=False))
        t0 = time.perf_counter()
        gr.run_basegrid()
        print("base grid: %.2f s" % (time.perf_counter() - t0), flush=True)
        for k in range(fills):
            t1 = time.perf_counter()
            gr.fill_in_grid()
            print("fill-in %d -> %s: %.2f s (cumulative %.2f)" % (k + 1, gr.grids[-1]['Kx'].shape,
                  time.perf_counter() - t1, time.perf_counter() - t0), flush=True)


if __name__ == "__main__":
    main()
