import sys; sys.path.insert(0,'/root/repo')
import numpy as np, torch
from mddatasetbuilder_b200.engine import Engine
eng=Engine(0)
def run(X, C, name):
    rows=len(X)
    lab,nre,dg=eng.kmeans_assign_tc(torch.as_tensor(X,device='cuda'), torch.as_tensor(C,device='cuda'), diag=True)
    dg=dg.cpu().numpy(); b1=dg[:rows].astype(np.float64); b2=dg[rows:2*rows]; i1=dg[2*rows:3*rows].view(np.int32); xn2=dg[3*rows:4*rows]; cmax2=dg[4*rows]
    mu=C.mean(0); Xc=X-mu; Cc=C-mu
    # the score the kernel maximises: x.c - ||c||^2/2 (centred); i1 is the first of 8 columns (stride 4) holding the best
    cand=np.minimum(i1[:,None]+4*np.arange(8)[None,:],len(C)-1)
    exact=((Xc[:,None,:]*Cc[cand]).sum(2)-0.5*(Cc[cand]**2).sum(2)).max(1)
    err=np.abs(b1-exact)
    norm=np.sqrt(xn2.astype(np.float64)*cmax2)
    sxc=(np.abs(Xc)*np.abs(Cc[np.minimum(i1,len(C)-1)])).sum(1)
    print(f"{name:28s} rows={rows} D={X.shape[1]} K={len(C)} recheck={nre} ({100*nre/rows:.2f}%) max|err|={err.max():.3e} max err/(|x||c|max)={np.max(err/norm):.3e} max err/sum|xc|={np.max(err/np.maximum(sxc,1e-300)):.3e}")
rng=np.random.default_rng(0)
for D,K in [(35,10000),(64,2500),(110,1000),(16,5000)]:
    X=rng.random((20000,D)); C=X[rng.choice(20000,K,replace=False)]+rng.normal(0,0.01,(K,D)); run(X,C,f"uniform D={D}")
proto=rng.random((500,40)); X=np.concatenate([p+rng.normal(0,1e-3,(40,40)) for p in proto]); C=np.concatenate([proto+rng.normal(0,1e-4,proto.shape)]*2); run(X,C,"clustered")
X=rng.normal(0,100,(20000,35))+500; C=X[:3000]+rng.normal(0,1,(3000,35)); run(X,C,"offset large magnitude")
