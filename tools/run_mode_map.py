import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from psitools_public_b200 import PSIMode, TanhSinhNoDeepCopy
n=int(sys.argv[1]) if len(sys.argv)>1 else 42
pm=PSIMode(3.0,[1e-8,1.0],[-2,2],[1e-8,1.0],tanhsinh_integrator=TanhSinhNoDeepCopy())
ctx=pm.psi_dispersion.ctx
ax=np.logspace(-1,3,n); kxg,kzg=np.meshgrid(ax,ax); N=n*n
roots,nc,st=ctx.mode_batch([pm.psi_dispersion.dist_id]*N, kxg.ravel(), kzg.ravel(), np.zeros(N), [[-2,2,1e-8,1.0]]*N, 20,1,100,1e-13,1e-4,[2]*N,[[]]*N, device_aaa=True)
print('ok', sum(len(r)>0 for r in roots))
