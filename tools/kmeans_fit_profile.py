"""Developer tool: wall time of the whole MiniBatchKMeans selection (init x3, mini-batch loop, final labels, picks)."""
import sys, time; sys.path.insert(0, '/root/repo')
import numpy as np, torch
from mddatasetbuilder_b200.engine import Engine
from mddatasetbuilder_b200.kmeans import clusterdatas, MiniBatchKMeansB200
eng = Engine(0)
rows, D, K = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000, 35, int(sys.argv[2]) if len(sys.argv) > 2 else 10000
g = torch.Generator(device='cuda'); g.manual_seed(1)
X = torch.rand((rows, D), dtype=torch.float64, device='cuda', generator=g)
for it in range(2):
    eng.profile(it == 1)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    km = MiniBatchKMeansB200(eng, n_clusters=K, init_size=min(3 * K, rows), n_init=3, seed=0)
    km.fit(X)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print('fit %d x %d, K=%d: %.2f s, n_steps %s' % (rows, D, K, dt, getattr(km, 'n_steps_', '?')))
pr = eng.profile_read()
tot = sum(ms for _, ms in pr.values())
print('GPU kernel time %.1f ms over %d launches' % (tot, sum(c for c, _ in pr.values())))
for k, (c, ms) in sorted(pr.items(), key=lambda kv: -kv[1][1])[:10]:
    print('  %-22s %7d launches %9.1f ms' % (k, c, ms))
