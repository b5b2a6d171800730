"""Tiny end-to-end run of every kernel (K0-K4) for compute-sanitizer:
    compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from psitools_public_b200 import PSIDispersion, TanhSinhNoDeepCopy, complex_roots_mpi, psi_mode_mpi  # noqa: E402

ts = TanhSinhNoDeepCopy()
pd = PSIDispersion(3.0, [1e-8, 1.0], tanhsinh_integrator=ts)
print("K0/K1:", pd.calculate(np.array([0.3 + 0.2j, -1.0048 + 1e-4j, 1.5 - 0.7j]), 10.0, 100.0))
z = [-2.0 + 2e-7j, 2.0 + 2e-7j, 2.0 + 2.0j, -2.0 + 2.0j]
items = [{'PSIDispersion.__init__': {'stokes_range': [1e-8, 1.0], 'dust_to_gas_ratio': 3.0,
                                     'tanhsinh_integrator': ts},
          'dispersion': {'wave_number_x': kx, 'wave_number_z': 10.0, 'viscous_alpha': 0.0},
          'z_list': z, 'count_roots': {'tol': 0.1}} for kx in (0.1, 10.0)]
print("K2:", complex_roots_mpi.MpiScheduler(time.time(), 1e9, verbose=False).run(items))
root = -1.004879704650626 + 0.0008946325671658642j
mitems = [{'__init__': {'stokes_range': [1e-8, 1.0], 'dust_to_gas_ratio': 3.0, 'real_range': [-2, 2],
                        'imag_range': [1e-8, 1], 'tanhsinh_integrator': ts},
           'calculate': {'wave_number_x': kx, 'wave_number_z': kz, 'guess_roots': g}, 'random_seed': 2}
          for kx, kz, g in ((0.1, 10.0, []), (1.0, 1.0, []), (0.1, 10.0, [root*(1 + 1e-3)]), (100.0, 1.0, []))]
for eng in ("device", "native"):
    res = psi_mode_mpi.MpiScheduler(time.time(), 1e9, verbose=False, engine=eng).run(mitems)
    print("K3/K4 (%s):" % eng, [list(np.round(v, 6)) for v in res.values()])

# K1 fast path (thread kernel + leftover kernels), tabulated density, K5 mirror, terminal-velocity map
w = (np.linspace(-2, 2, 96)[None, :] + 1j*np.linspace(-0.02, 0.5, 64)[:, None]).ravel()
print("K1 fast path:", np.abs(pd.calculate(w, 30.0, 30.0)).max())
from psitools_public_b200 import PSIGridRefiner, PSIModeTV, SizeDensity  # noqa: E402
sd = SizeDensity(lambda a: a**0.5*np.exp(-(np.log(a/0.01))**2/8.0), [1e-6, 0.5])
pt = PSIDispersion(2.0, [1e-6, 0.5], size_density=sd, tanhsinh_integrator=ts)
print("table:", pt.calculate(np.array([0.3 + 0.2j, 0.1 + 1e-3j]), 10.0, 50.0))
base = {'__init__': {'stokes_range': [1e-8, 1.0], 'dust_to_gas_ratio': 3.0, 'real_range': [-2.0, 2.0],
                     'imag_range': [1e-8, 1.0], 'n_sample': 20, 'max_zoom_domains': 1, 'tol': 1.0e-13,
                     'clean_tol': 1e-4, 'tanhsinh_integrator': ts},
        'calculate': {'wave_number_x': None, 'wave_number_z': None, 'guess_roots': []}, 'random_seed': 2}
gr = PSIGridRefiner('x', baseargs=base, nbase=(3, 3), reruns=1, krange=(1.0, 1.6), tv_guesses=True)
gr.run_basegrid()
gr.fill_in_grid()
print("K5:", gr.sweep_log)
print("tv:", PSIModeTV(3.0, 0.1).find_roots_map(1e-8, np.logspace(0, 2, 5), np.logspace(0, 2, 5))[1])
