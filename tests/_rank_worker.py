"""Worker for tests/test_multi_rank.py: one rank of a world_size-2 gloo job (CPU).  Uses the
lane-emulation backend for D(omega); exercises the sharding + all-gather path of the schedulers."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch.distributed as dist
    dist.init_process_group("gloo")
    import hostsim_util as hs
    from psitools_public_b200 import _device, _dist, TanhSinhNoDeepCopy, complex_roots_mpi, psi_mode_mpi
    cache = {}
    _device.get_context = lambda integ, device=None: cache.setdefault("c", hs.HostsimContext(integ))
    from test_host_logic import count_arglist, grid_arglist
    ts = TanhSinhNoDeepCopy()
    out = {"rank": _dist.rank(), "world": _dist.world_size(), "mine": _dist.my_items(9)}
    ms = psi_mode_mpi.MpiScheduler(time.time(), 3600, verbose=False)
    allres = ms.run_all(grid_arglist(3, ts))
    out["mode_items_local"] = ms.last_stats["items"]
    out["mode"] = {str(k): [[z.real, z.imag] for z in v] for k, v in allres.items()}
    out["run_is_none_on_workers"] = (ms.run(grid_arglist(3, ts)[:2]) is None) == (_dist.rank() != 0)
    d = psi_mode_mpi.DispersionRelationMpiScheduler(time.time(), 3600, verbose=False)
    items = [{'__init__': {'stokes_range': [1e-8, 1.0], 'dust_to_gas_ratio': 3.0, 'real_range': [-2, 2],
                           'imag_range': [-2, 2], 'tanhsinh_integrator': ts},
              'dispersion': {'w': w, 'wave_number_x': 1e2, 'wave_number_z': 1e1, 'viscous_alpha': 0.0}}
             for w in (-2 - 2j, 2 - 2j, -2 + 2j)]
    dv = d.run_all(items)
    out["disp"] = {str(k): [complex(v).real, complex(v).imag] for k, v in dv.items()}
    c = complex_roots_mpi.MpiScheduler(time.time(), 3600, verbose=False).run_all(count_arglist(ts)[:4])
    out["counts"] = {str(k): v for k, v in c.items()}
    # grid refinement through the array path (MpiScheduler.run_arrays: rank-strided slices of the item
    # arrays and of the CSR guess lists, dense all-gather, replicated RootStore on every rank)
    from psitools_public_b200 import PSIGridRefiner
    base = {'__init__': {'stokes_range': [1e-8, 1.0], 'dust_to_gas_ratio': 3.0, 'real_range': [-2.0, 2.0],
                         'imag_range': [1e-8, 1.0], 'n_sample': 20, 'max_zoom_domains': 1, 'tol': 1.0e-13,
                         'clean_tol': 1e-4, 'tanhsinh_integrator': ts},
            'calculate': {'wave_number_x': None, 'wave_number_z': None, 'guess_roots': []}, 'random_seed': 2}
    gr = PSIGridRefiner('x', baseargs=base, nbase=(3, 3), reruns=1, krange=(1.0, 1.6),
                        scheduler=psi_mode_mpi.MpiScheduler(time.time(), 3600, verbose=False))
    assert gr._array_path()
    gr.run_basegrid()
    gr.fill_in_grid()
    g = gr.grids[-1]
    out["refine"] = {"log": [list(x) for x in gr.sweep_log], "runs": g['runs'].tolist(),
                     "roots": {str(k): [[z.real, z.imag] for z in g['results'][k]] for k in g['results']}}
    with open(os.path.join(sys.argv[1], "rank%d.json" % _dist.rank()), "w") as fh:
        json.dump(out, fh)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
