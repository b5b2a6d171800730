"""Compare the device (K4 pipeline) and native (host LAPACK AAA) engines on an n x n PSIMode map:
de-duplicated root sets per pixel and wall time.
usage: python tools/compare_engines.py [n] [mrn|bump|tail|visc|mrn_small]"""
import os
import sys
import time

import numpy as np

_R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, _R)
sys.path.insert(0, os.path.join(_R, 'tests'))
from psitools_public_b200 import PowerBump, PowerBumpTail, PSIMode, SizeDensity, TanhSinhNoDeepCopy  # noqa: E402
from test_grid_refine import distinct  # noqa: E402
from test_host_logic import same_root_sets  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
kind = sys.argv[2] if len(sys.argv) > 2 else "mrn"
ts = TanhSinhNoDeepCopy()
alpha = 0.0
if kind in ("bump", "tail"):
    cls = PowerBump if kind == "bump" else PowerBumpTail
    pb = cls(amin=1e-8, aP=0.1, aL=2.0/3.0*0.1, aR=0.156, bumpfac=2.0)
    sd = SizeDensity(pb.sigma0, [1e-8, 0.156])
    sd.poles = [pb.get_discontinuity()]
    pm = PSIMode(10.0, [1e-8, 0.156], [-2, 2], [1e-7, 1.0], size_density=sd, tanhsinh_integrator=ts)
elif kind == "visc":
    pm = PSIMode(3.0, [1e-8, 0.1], [-2, 2], [1e-8, 1.0], tanhsinh_integrator=ts)
    alpha = 1e-6
elif kind == "mrn_small":
    pm = PSIMode(3.0, [1e-8, 1e-2], [-2, 2], [1e-8, 1.0], tanhsinh_integrator=ts)
else:
    pm = PSIMode(3.0, [1e-8, 1.0], [-2, 2], [1e-8, 1.0], tanhsinh_integrator=ts)
ctx = pm.psi_dispersion.ctx
ax = np.logspace(-1, 3, n)
kxg, kzg = np.meshgrid(ax, ax)
N = n*n
dom = [[pm.domain.xmin, pm.domain.xmax, pm.domain.ymin, pm.domain.ymax]]*N
res = {}
for eng in ('device', 'native', 'device'):
    t = time.time()
    roots, nc, st = ctx.mode_batch([pm.psi_dispersion.dist_id]*N, kxg.ravel(), kzg.ravel(), np.full(N, alpha), dom,
                                   20, 1, 100, 1e-13, 1e-4, [2]*N, [[]]*N, device_aaa=(eng == 'device'))
    dt = time.time() - t
    print('%s %s grid %dx%d: %.2f s, %.1f modes/s, evals %d (%.1f/px), pixels with roots %d, launches %s'
          % (kind, eng, n, n, dt, N/dt, st['evals'], st['evals']/N, sum(len(r) > 0 for r in roots),
             st.get('launches')), flush=True)
    res[eng] = roots
bad = [i for i in range(N) if not same_root_sets(distinct(res['device'][i]), distinct(res['native'][i]))]
print('pixels whose de-duplicated root sets differ (device vs native):', len(bad), 'of', N)
for i in bad[:10]:
    print('  Kx %.4g Kz %.4g device %s native %s' % (kxg.ravel()[i], kzg.ravel()[i], res['device'][i], res['native'][i]))
