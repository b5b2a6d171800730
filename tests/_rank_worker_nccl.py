"""Worker for tests/test_multi_rank.py::test_two_gpu_device_records: one rank of a world_size-2 NCCL job on real
GPUs.  The growth-rate map and the refinement campaign go through the device-to-device exchange (packed records
out of the library's buffer, one ncclAllGather); results are written per rank for the parent test to compare with
a single-GPU run."""
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def campaign(ms_factory):
    from psitools_public_b200 import PSIGridRefiner, TanhSinhNoDeepCopy
    ts = TanhSinhNoDeepCopy()
    from test_host_logic import grid_arglist
    out = {}
    ms = ms_factory()
    t0 = time.time()
    res = ms.run_all(grid_arglist(24, ts))
    out["map_seconds"] = time.time() - t0
    out["exchange"] = ms.last_stats.get("exchange")
    out["map_hash"] = hashlib.sha256(b"".join(np.asarray(res[i], dtype=complex).tobytes()
                                              for i in range(len(res)))).hexdigest()
    out["map_with_roots"] = int(sum(len(v) > 0 for v in res.values()))
    r0 = ms.run(grid_arglist(5, ts))
    out["run_is_none"] = r0 is None
    base = {'__init__': {'stokes_range': [1e-8, 1.0], 'dust_to_gas_ratio': 3.0, 'real_range': [-2.0, 2.0],
                         'imag_range': [1e-8, 1.0], 'n_sample': 20, 'max_zoom_domains': 1, 'tol': 1.0e-13,
                         'clean_tol': 1e-4, 'tanhsinh_integrator': ts},
            'calculate': {'wave_number_x': None, 'wave_number_z': None, 'guess_roots': []}, 'random_seed': 2}
    gr = PSIGridRefiner('x', baseargs=base, nbase=(6, 6), reruns=1, krange=(0.5, 2.0), scheduler=ms_factory())
    gr.run_basegrid()
    gr.fill_in_grid()
    g = gr.grids[-1]
    out["refine_exchange"] = gr.ms.last_stats.get("exchange")
    out["refine_log"] = [list(x) for x in gr.sweep_log]
    out["refine_runs"] = g['runs'].tolist()
    out["refine_hash"] = hashlib.sha256(b"".join(np.asarray(g['results'][k], dtype=complex).tobytes()
                                                 for k in g['results'])).hexdigest()
    return out


def main():
    import torch
    import torch.distributed as dist
    from psitools_public_b200 import _dist, psi_mode_mpi
    multi = "RANK" in os.environ
    if multi:
        local = int(os.environ["LOCAL_RANK"])
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    out = campaign(lambda: psi_mode_mpi.MpiScheduler(time.time(), 3600, verbose=False))
    out["rank"], out["world"] = _dist.rank(), _dist.world_size()
    with open(os.path.join(sys.argv[1], "rank%d_of%d.json" % (_dist.rank(), _dist.world_size())), "w") as fh:
        json.dump(out, fh)
    if multi:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
