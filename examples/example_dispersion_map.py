#!/usr/bin/env python
"""D(omega) on a dense complex-frequency grid at fixed (Kx, Kz) -- the workload of the reference's
psitools_examples/example_PSIModeMPIDispersion.py (:15-88), through DispersionRelationMpiScheduler.

    python examples/example_dispersion_map.py [N]            (1 GPU)
    torchrun --nproc-per-node 8 examples/example_dispersion_map.py 1024

Writes <batchname>.hdf5 with the reference's dataset names (Rgrid, Igrid, dispersion_real,
dispersion_imag) when h5py is importable, else <batchname>.npz with the same names."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from psitools_public_b200 import _dist, psi_mode_mpi, tanhsinh  # noqa: E402


def main(n=256, batchname="dispersion_map"):
    if "RANK" in os.environ:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
        dist.init_process_group("nccl")
    real_range, imag_range = [-7.5, 7.5], [-12.5, 2.5]
    rgrid, igrid = np.meshgrid(np.linspace(*real_range, n), np.linspace(*imag_range, n))
    init = {'stokes_range': [1e-8, 1e-1], 'dust_to_gas_ratio': 3.0, 'real_range': real_range,
            'imag_range': imag_range, 'tanhsinh_integrator': tanhsinh.TanhSinhNoDeepCopy()}
    arglist = [{'__init__': init,
                'dispersion': {'w': complex(a, b), 'wave_number_x': 50.0, 'wave_number_z': 100.0,
                               'viscous_alpha': 0.0}} for a, b in zip(rgrid.ravel(), igrid.ravel())]
    t0 = time.time()
    ms = psi_mode_mpi.DispersionRelationMpiScheduler(t0, 70*60*60, verbose=False)
    finished = ms.run(arglist)
    if finished is None:                      # not rank 0
        return
    vals = np.array([finished[i] for i in range(len(arglist))]).reshape(rgrid.shape)
    print("%d D(omega) values in %.2f s on %d rank(s)" % (vals.size, time.time() - t0, _dist.world_size()))
    data = dict(Rgrid=rgrid, Igrid=igrid, dispersion_real=vals.real, dispersion_imag=vals.imag)
    try:
        import h5py
        with h5py.File(batchname + ".hdf5", "w") as h5f:
            grp = h5f.create_group(batchname)
            for k, v in init.items():
                if k != 'tanhsinh_integrator':
                    grp.attrs[k] = v
            for k, v in data.items():
                grp.create_dataset(k, data=v)
    except ImportError:
        np.savez(batchname + ".npz", **data)


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 256)
