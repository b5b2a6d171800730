#!/usr/bin/env python
"""Growth-rate map with terminal-velocity guesses: the mode of the terminal-velocity approximation
(reference psitools/terminalvelocitysolver.py, PSIModeTV) is computed for the whole (Kx, Kz) map in one
kernel (PSIModeTV.find_roots_map -> psi_tv_roots) and handed to the PSIMode root searches of the map as
``guess_roots`` where it exists, grows and lies inside the search domain.

    python examples/example_tv_guesses.py [N]

Prints how many pixels have a growing root with and without the guesses."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from psitools_public_b200 import PSIModeTV, psi_mode_mpi, tanhsinh  # noqa: E402


def main(n=64):
    mu, stokes_range = 3.0, [1e-8, 0.1]
    real_range, imag_range = [-2.0, 2.0], [1e-8, 1.0]
    axis = np.logspace(-1, 3, n)
    kxg, kzg = np.meshgrid(axis, axis)
    t0 = time.time()
    w_tv, status = PSIModeTV(mu, stokes_range[1]).find_roots_map(stokes_range[0], kxg, kzg)
    ok = (status == 0) & (w_tv.imag > imag_range[0]) & (w_tv.imag < imag_range[1]) & \
        (w_tv.real > real_range[0]) & (w_tv.real < real_range[1])
    print("terminal-velocity map %dx%d in %.3f s: a usable mode on %d pixels" % (n, n, time.time() - t0, ok.sum()))
    init = {'stokes_range': stokes_range, 'dust_to_gas_ratio': mu, 'real_range': real_range,
            'imag_range': imag_range, 'n_sample': 20, 'max_zoom_domains': 1,
            'tanhsinh_integrator': tanhsinh.TanhSinhNoDeepCopy()}
    for label, guided in (("without guesses", False), ("with terminal-velocity guesses", True)):
        arglist = [{'__init__': init,
                    'calculate': {'wave_number_x': float(a), 'wave_number_z': float(b),
                                  'guess_roots': [complex(g)] if (guided and use) else []},
                    'random_seed': 2}
                   for a, b, g, use in zip(kxg.ravel(), kzg.ravel(), w_tv.ravel(), ok.ravel())]
        t0 = time.time()
        res = psi_mode_mpi.MpiScheduler(t0, 70*60*60, verbose=False).run(arglist)
        growing = sum(1 for r in res.values() if len(r) and np.max(np.imag(r)) > 0)
        print("%-32s %d searches in %.2f s, pixels with a growing root: %d" % (label, len(arglist), time.time() - t0,
                                                                              growing))


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 64)
