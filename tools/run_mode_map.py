"""The n x n PSIMode map of BASELINE config 3 through the K4 pipeline (device AAA), with the stage times and a
sha256 of all roots -- for ncu captures and for A/B runs of two builds of the library:
    python tools/run_mode_map.py 256 [path/to/other/libpsi_b200.so]"""
import hashlib
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from psitools_public_b200 import PSIMode, TanhSinhNoDeepCopy, _device  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 42
if len(sys.argv) > 2:
    _device.LIB_PATH = os.path.abspath(sys.argv[2])
pm = PSIMode(3.0, [1e-8, 1.0], [-2, 2], [1e-8, 1.0], tanhsinh_integrator=TanhSinhNoDeepCopy())
ctx = pm.psi_dispersion.ctx
ax = np.logspace(-1, 3, n)
kxg, kzg = np.meshgrid(ax, ax)
N = n*n
for rep in range(int(os.environ.get("PSI_REPS", "1"))):
    t0 = time.perf_counter()
    roots, nc, st = ctx.mode_batch([pm.psi_dispersion.dist_id]*N, kxg.ravel(), kzg.ravel(), np.zeros(N),
                                   [[-2, 2, 1e-8, 1.0]]*N, 20, 1, 100, 1e-13, 1e-4, [2]*N, [[]]*N, device_aaa=True)
    dt = time.perf_counter() - t0
    h = hashlib.sha256(b''.join(np.asarray(r, dtype=complex).tobytes() for r in roots)).hexdigest()[:16]
    if os.environ.get("PSI_SAVE"):
        m = np.full((N, 8), np.nan, dtype=complex)
        for i, r in enumerate(roots):
            m[i, :min(len(r), 8)] = np.asarray(r)[:8]
        np.save(os.environ["PSI_SAVE"], m)
    print('ok', sum(len(r) > 0 for r in roots), 'pixels with roots; %.3f s;' % dt, 'stages',
          {k: round(v, 4) for k, v in (st.get('stage_s') or {}).items()}, 'roots sha256', h, 'lib', _device.LIB_PATH)
