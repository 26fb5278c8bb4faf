"""One tensor-core assignment call on bench-shaped data (for ncu captures); prints the per-kernel times."""
import sys; sys.path.insert(0, '/root/repo')
import numpy as np, torch
from mddatasetbuilder_b200.engine import Engine
eng = Engine(0)
g = torch.Generator(device='cuda'); g.manual_seed(1)
X = torch.rand((262144, 35), dtype=torch.float64, device='cuda', generator=g)
C = X[torch.randperm(262144, device='cuda', generator=g)[:10000]].contiguous() + 0.01
for it in range(3):
    eng.profile(it == 2)
    lab, nre = eng.kmeans_assign_tc(X, C)
torch.cuda.synchronize()
print('uniform', nre, {k: (c, round(ms, 3)) for k, (c, ms) in eng.profile_read().items()})
