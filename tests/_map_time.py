import sys, time, os, numpy as np
sys.path.insert(0, '/root/repo')
from psitools_public_b200 import PSIMode, TanhSinhNoDeepCopy
n=int(sys.argv[1]) if len(sys.argv)>1 else 64
pm=PSIMode(3.0,[1e-8,1.0],[-2,2],[1e-8,1.0],tanhsinh_integrator=TanhSinhNoDeepCopy())
ctx=pm.psi_dispersion.ctx
ax=np.logspace(-1,3,n); kxg,kzg=np.meshgrid(ax,ax); N=n*n
for rep in range(2):
    t=time.time()
    roots,nc,st=ctx.mode_batch([pm.psi_dispersion.dist_id]*N, kxg.ravel(), kzg.ravel(), np.zeros(N), [[-2,2,1e-8,1.0]]*N, 20,1,100,1e-13,1e-4,[2]*N,[[]]*N)
    dt=time.time()-t
    print('grid %dx%d: %.2f s, %.1f modes/s, passes %d evals %d (%.1f/px), pixels with roots %d, cores %d'%(n,n,dt,N/dt,st['passes'],st['evals'],st['evals']/N,sum(len(r)>0 for r in roots), os.cpu_count()))
