#!/usr/bin/env python
"""Number of growing roots per (Kx, Kz) pixel by the argument principle -- the workload of the
reference's psitools_examples/example_psicountgrid.py (:12-83), through complex_roots_mpi.MpiScheduler
(kernel K2: one warp per contour).

    python examples/example_count_grid.py [N]

Writes <batchname>.hdf5 (datasets Kxgrid, Kzgrid, root_count) or <batchname>.npz without h5py."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from psitools_public_b200 import complex_roots_mpi, tanhsinh  # noqa: E402


def main(n=32, batchname="count_grid"):
    if "RANK" in os.environ:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
        dist.init_process_group("nccl")
    domain_real, domain_imag = [-2.0, 2.0], [2e-7, 2.0]
    z_list = [domain_real[0] + 1j*domain_imag[0], domain_real[1] + 1j*domain_imag[0],
              domain_real[1] + 1j*domain_imag[1], domain_real[0] + 1j*domain_imag[1]]
    axis = np.logspace(-1, 3, n)
    kxg, kzg = np.meshgrid(axis, axis)
    init = {'stokes_range': [1e-8, 1.0], 'dust_to_gas_ratio': 3.0, 'size_distribution_power': 3.5,
            'tanhsinh_integrator': tanhsinh.TanhSinhNoDeepCopy()}
    arglist = [{'PSIDispersion.__init__': init,
                'dispersion': {'wave_number_x': float(a), 'wave_number_z': float(b), 'viscous_alpha': 0.0},
                'z_list': z_list, 'count_roots': {'tol': 0.1}} for a, b in zip(kxg.ravel(), kzg.ravel())]
    t0 = time.time()
    finished = complex_roots_mpi.MpiScheduler(t0, 70*60*60, verbose=False).run(arglist)
    if finished is None:
        return
    counts = np.array([finished[i] for i in range(len(arglist))]).reshape(kxg.shape)
    print("%d contours in %.2f s; pixels with a growing root: %d" % (counts.size, time.time() - t0,
                                                                   int(np.sum(counts > 0))))
    data = dict(Kxgrid=kxg, Kzgrid=kzg, root_count=counts)
    try:
        import h5py
        with h5py.File(batchname + ".hdf5", "w") as h5f:
            grp = h5f.create_group(batchname)
            for k, v in data.items():
                grp.create_dataset(k, data=v)
    except ImportError:
        np.savez(batchname + ".npz", **data)


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 32)
