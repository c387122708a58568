import sys; sys.path.insert(0,'/root/repo')
import numpy as np, torch
from sparsex_b200 import CsxMatrix
d=np.load('/root/repo/tmp_debug/case_s3.npz'); rp,ci,va,n=d['rp'],d['ci'],d['va'],int(d['n'])
x=np.random.default_rng(3).uniform(-1,1,n)
rows=np.repeat(np.arange(n),np.diff(rp)); yref=np.zeros(n); np.add.at(yref,rows,va*x[ci])
XF=sys.argv[1]
for sym in ['true']:
  for nt in [1,2,3,4,5,6,8]:
    A=CsxMatrix.tune_csr(rp,ci,va,n,n,{'spx.preproc.xform':XF,'spx.matrix.symmetric':sym,'spx.rt.nr_threads':nt}).upload(0)
    dx=torch.from_numpy(x).cuda(); dy=torch.zeros(n,dtype=torch.float64,device='cuda')
    for rep in range(3):
        A.spmv(1.0,dx,dy); torch.cuda.synchronize()
        y=dy.cpu().numpy(); bad=np.nonzero(np.abs(y-yref)>1e-9)[0]
        print(sym,nt,rep,'maxerr',np.abs(y-yref).max(),'bad',bad[:12], [ (A.partition(p).row_start, len(A.partition(p).dvalues)) for p in range(A.nparts)] if rep==0 else '')
