"""D(omega) on the config-2 grid with the library named by PSI_B200_LIB; saves D and the node count."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from psitools_public_b200 import PSIDispersion, TanhSinhNoDeepCopy
out = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
pd = PSIDispersion(3.0, [1e-8, 0.1], tanhsinh_integrator=TanhSinhNoDeepCopy())
re, im = np.meshgrid(np.linspace(-7.5, 7.5, n), np.linspace(-12.5, 2.5, n))
w = (re + 1j*im).ravel()
pd.ctx.disp_eval(pd.dist_id, 50.0, 100.0, 0.0, w[:1000])
n0 = pd.ctx.node_evals
t = time.time()
D = pd.ctx.disp_eval(pd.dist_id, 50.0, 100.0, 0.0, w)
dt = time.time() - t
print(out, "seconds %.4f" % dt, "nodes/point %.0f" % ((pd.ctx.node_evals - n0)/len(w)))
np.save(out, D)
