"""Developer tool: clock64 stamps of CTA 0 of k_assign_tc (MDB_TC_TRACE=1), per centroid tile.
Columns: MMA warp [loop top, t_empty ok, b_full ok, issued] and one epilogue warp [at wait, t_full ok, TMEM released, tile done]."""
import os, sys; sys.path.insert(0, '/root/repo')
os.environ['MDB_TC_TRACE'] = '1'
import numpy as np, torch
from mddatasetbuilder_b200.engine import Engine
eng = Engine(0)
g = torch.Generator(device='cuda'); g.manual_seed(1)
X = torch.rand((262144, 35), dtype=torch.float64, device='cuda', generator=g)
C = X[torch.randperm(262144, device='cuda', generator=g)[:10000]].contiguous() + 0.01
for it in range(2):
    lab, nre, dg = eng.kmeans_assign_tc(X, C, diag=True)
torch.cuda.synchronize()
tr = dg[:79 * 16].view(torch.int64).cpu().numpy().reshape(79, 8)
t0 = tr[0, 0]
for t in list(range(0, 12)) + list(range(70, 79)):
    print(t, ' '.join('%7d' % (v - t0) for v in tr[t, :8]))
print('cycles per tile (steady):', (tr[70, 0] - tr[10, 0]) / 60.0)
