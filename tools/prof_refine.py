"""Where the analysis stage of a PSIGridRefiner campaign (BASELINE config 4) spends its cycles, sweep by sweep,
from a -DPSI_PROF build (tools/build_prof.sh):
    PSI_B200_LIB=_exp/libpsi_prof.so python tools/prof_refine.py [fills]
Per sweep: items, analysis seconds, summed warp-cycles in the analysis and in its QR factorisations, the share of
QR cycles spent at more than 48 support points, the largest support count and the LONGEST single analysis."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench_extra  # noqa: E402
from psitools_public_b200 import _device, psi_mode_mpi  # noqa: E402

fills = int(sys.argv[1]) if len(sys.argv) > 1 else 6
lib = _device.load_library()
lib.psi_prof_read.argtypes = [C.c_void_p]
buf = np.zeros(32, dtype=np.uint64)
inner = psi_mode_mpi.MpiScheduler.run_arrays
rows = []
best = [0]
phases = []


def wrapped(self, *a, **k):
    lib.psi_prof_read(buf.ctypes.data_as(C.c_void_p))          # read + reset
    out = inner(self, *a, **k)
    lib.psi_prof_read(buf.ctypes.data_as(C.c_void_p))
    st = self.last_stats
    cyc, idx = int(buf[22]) >> 22, int(buf[22]) & 0x3fffff
    buf[22] = cyc
    if cyc > best[0]:                           # keep the inputs of the slowest search for a post-mortem
        best[0] = cyc
        kx, kz, seeds = a[1], a[2], a[3]
        g = k.get("guesses") if "guesses" in k else (a[4] if len(a) > 4 else None)
        gl = np.zeros(0, complex) if g is None else np.asarray(g[0])[int(g[1][idx]):int(g[1][idx + 1])]
        os.makedirs("gpurun_out", exist_ok=True)
        np.savez("gpurun_out/slowest_search.npz", kx=kx[idx], kz=kz[idx], seed=np.broadcast_to(seeds, (len(kx),))[idx],
                 guesses=gl, cycles=cyc, items=len(kx), idx=idx)
    phases.append((len(a[1]), buf.copy()))
    rows.append((len(a[1]), (st.get("stage_s") or {}).get("analyze", 0.0), int(buf[12]), int(buf[3]), int(buf[20]),
                 int(buf[21]), int(buf[22])))
    return out


psi_mode_mpi.MpiScheduler.run_arrays = wrapped
os.environ.pop("PSI_BENCH_PER_SWEEP", None)
bench_extra.refine(16, fills, 3)
print("%9s %9s %12s %12s %8s %6s %12s" % ("items", "analyze_s", "an Mcycles", "QR Mcycles", "QR m>48", "max m", "longest ms"))
for n, s, tot, qr, big, mm, longest in rows:
    print("%9d %9.4f %12.1f %12.1f %7.1f%% %6d %12.2f" % (n, s, tot/1e6, qr/1e6, 100.0*big/max(qr, 1), mm,
                                                         longest/1.9e6))

NAMES = {0: "index lists", 1: "memo lookup", 2: "Loewner build", 3: "QR", 4: "inverse iteration", 5: "residuals",
         6: "memo insert", 7: "Aberth (poles)", 8: "Aberth (zeros)", 9: "zeros: secant on r(z)", 10: "zeros: sort/unique",
         11: "cleanup: residues"}
for label, sel in (("whole campaign", lambda n: True), ("sweeps with guesses > 10k items", lambda n: n in (12944, 26080))):
    tot = sum(b for n, b in phases if sel(n))
    print(label, "analysis Mcycles %.0f" % (tot[12]/1e6))
    for k in sorted(NAMES):
        print("   %-24s %6.1f %%" % (NAMES[k], 100.0*tot[k]/max(float(tot[12]), 1.0)))
