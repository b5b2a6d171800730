"""Cycles per phase of the AAA analysis stage (kernel mode_analyze_kernel and its successors) on the n x n map of
BASELINE config 3, from a -DPSI_PROF build of the library (clock64 on lane 0 of every warp, summed):
    nvcc ... -DPSI_PROF -o _exp/libpsi_prof.so ...        (tools/build_prof.sh)
    PSI_B200_LIB=_exp/libpsi_prof.so python tools/prof_analysis.py 256"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from psitools_public_b200 import PSIMode, TanhSinhNoDeepCopy, _device  # noqa: E402

NAMES = {0: "index lists", 1: "memo lookup", 2: "Loewner build", 3: "QR", 4: "inverse iteration", 5: "residuals",
         6: "memo insert", 7: "Aberth (poles)", 8: "Aberth (zeros)", 9: "zeros: secant on r(z)", 10: "zeros: sort/unique",
         11: "cleanup: residues", 12: "mode_analyze total", 13: "  first aaa_calculate", 14: "  clean_tol halving loop",
         15: "fit: greedy", 16: "fit: after cleanup", 17: "phase G kernel", 18: "phase C kernel"}

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
pm = PSIMode(3.0, [1e-8, 1.0], [-2, 2], [1e-8, 1.0], tanhsinh_integrator=TanhSinhNoDeepCopy())
ctx = pm.psi_dispersion.ctx
lib = ctx.lib
lib.psi_prof_read.argtypes = [C.c_void_p]
ax = np.logspace(-1, 3, n)
kxg, kzg = np.meshgrid(ax, ax)
N = n*n
buf = np.zeros(32, dtype=np.uint64)
for max_zoom in (0, 1):
    for rep in range(2):
        lib.psi_prof_read(buf.ctypes.data_as(C.c_void_p))
        roots, nc, st = ctx.mode_batch([pm.psi_dispersion.dist_id]*N, kxg.ravel(), kzg.ravel(), np.zeros(N),
                                       [[-2, 2, 1e-8, 1.0]]*N, 20, max_zoom, 100, 1e-13, 1e-4, [2]*N, [[]]*N,
                                       device_aaa=True)
    lib.psi_prof_read(buf.ctypes.data_as(C.c_void_p))
    tot = float(buf[12]) or 1.0
    print("max_zoom %d: %d searches, stage seconds %s" % (max_zoom, N, {k: round(v, 4) for k, v in st["stage_s"].items()}))
    for k in sorted(NAMES):
        if buf[k]:
            print("   %-28s %8.1f Mcycles/1k items   %5.1f %%" % (NAMES[k], buf[k]/1e6/(N/1e3), 100*buf[k]/tot))
