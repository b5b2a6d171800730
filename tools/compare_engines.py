"""Compare the device (K4 pipeline) and native (host LAPACK AAA) engines on an n x n PSIMode map:
de-duplicated root sets per pixel and wall time.  usage: python tools/compare_engines.py [n]"""
import sys, time, os, numpy as np
import os; _R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, _R); sys.path.insert(0, os.path.join(_R, 'tests'))
from psitools_public_b200 import PSIMode, TanhSinhNoDeepCopy
from test_host_logic import same_root_sets
from test_grid_refine import distinct
n=int(sys.argv[1]) if len(sys.argv)>1 else 64
pm=PSIMode(3.0,[1e-8,1.0],[-2,2],[1e-8,1.0],tanhsinh_integrator=TanhSinhNoDeepCopy())
ctx=pm.psi_dispersion.ctx
ax=np.logspace(-1,3,n); kxg,kzg=np.meshgrid(ax,ax); N=n*n
res={}
for eng in ('device','native','device'):
    t=time.time()
    roots,nc,st=ctx.mode_batch([pm.psi_dispersion.dist_id]*N, kxg.ravel(), kzg.ravel(), np.zeros(N), [[-2,2,1e-8,1.0]]*N, 20,1,100,1e-13,1e-4,[2]*N,[[]]*N, device_aaa=(eng=='device'))
    dt=time.time()-t
    print('%s grid %dx%d: %.2f s, %.1f modes/s, evals %d (%.1f/px), pixels with roots %d, launches %s'%(eng,n,n,dt,N/dt,st['evals'],st['evals']/N,sum(len(r)>0 for r in roots), st.get('launches')), flush=True)
    res[eng]=roots
bad=[i for i in range(N) if not same_root_sets(distinct(res['device'][i]), distinct(res['native'][i]))]
print('pixels whose de-duplicated root sets differ (device vs native):', len(bad), 'of', N)
for i in bad[:10]: print('  Kx %.4g Kz %.4g device %s native %s'%(kxg.ravel()[i], kzg.ravel()[i], res['device'][i], res['native'][i]))
