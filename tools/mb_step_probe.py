"""Developer probe: cost of the pieces of one mini-batch step (1024 rows, K = 10,000, D = 70)."""
import sys, time; sys.path.insert(0, '/root/repo')
import numpy as np, torch
from mddatasetbuilder_b200.engine import Engine
eng = Engine(0, ("C", "H", "O"))
dev = eng.device
n, D, K, B = 1 << 20, 70, 10000, 1024
X = torch.rand((n, D), dtype=torch.float64, device=dev)
C = X[torch.randperm(n, device=dev)[:K]].contiguous()
ws = torch.ones(K, dtype=torch.float64, device=dev)
idx = torch.randint(0, n, (B,), dtype=torch.int32, device=dev)
def t(fn, reps=50):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps * 1e6
print('assign fp64 + inertia (sync): %.0f us' % t(lambda: eng.kmeans_assign(X, C, rowidx=idx, want_inertia=True)))
print('assign fp64 no inertia:       %.0f us' % t(lambda: eng.kmeans_assign(X, C, rowidx=idx)))
print('assign tc (sync):             %.0f us' % t(lambda: eng.kmeans_assign_tc(X, C, rowidx=idx)))
lab, _, _ = eng.kmeans_assign(X, C, rowidx=idx)
labt, nre = eng.kmeans_assign_tc(X, C, rowidx=idx)
print('labels equal', bool(torch.equal(lab, labt)), 'rechecked', nre)
print('inertia kernel (sync):        %.0f us' % t(lambda: eng.kmeans_inertia(X, C, lab, rowidx=idx)))
print('update exact (1 block):       %.0f us' % t(lambda: eng.kmeans_update_exact(X, lab, C, ws, rowidx=idx)))
buf = (torch.empty((K, D), dtype=torch.float64, device=dev), torch.empty(K, dtype=torch.float64, device=dev))
print('partial sums + apply:         %.0f us' % t(lambda: (eng.kmeans_partial_sums(X, lab, K, rowidx=idx, out=buf), eng.kmeans_apply_update(C, ws, *buf))))
print('h2d of 1024 indices:          %.0f us' % t(lambda: torch.as_tensor(np.random.randint(0, n, B), dtype=torch.int32, device=dev)))
