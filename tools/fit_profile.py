"""Developer probe: where the wall time of MiniBatchKMeansB200.run_loop goes (cProfile over 600 steps)."""
import sys, time, cProfile, pstats; sys.path.insert(0, '/root/repo')
import numpy as np, torch
from mddatasetbuilder_b200.engine import Engine
from mddatasetbuilder_b200.kmeans import MiniBatchKMeansB200
eng = Engine(0, ("C", "H", "O"))
n, D, K = 1 << 21, 70, 10000
X = torch.rand((n, D), dtype=torch.float64, device=eng.device)
km = MiniBatchKMeansB200(eng, n_clusters=K, init_size=3 * K, n_init=1, seed=0, max_no_improvement=None, max_iter=1)
state = {"rs": np.random.RandomState(0), "centers": X[:K].clone(), "batch_size": 1024}
km.max_iter = 1
orig = km.run_loop
import types
# bound the number of steps: n_steps = max_iter * n // batch -> 2048; stop after 600 via max_no_improvement hack
class Stop(Exception): pass
cnt = [0]
real = km._mini_batch_step
def step(*a, **k):
    cnt[0] += 1
    if cnt[0] > 600: raise Stop
    return real(*a, **k)
km._mini_batch_step = step
torch.cuda.synchronize(); t0 = time.perf_counter()
pr = cProfile.Profile(); pr.enable()
try: km.run_loop(X, state)
except Stop: pass
pr.disable(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print('600 steps: %.1f ms total, %.0f us per step' % (dt * 1e3, dt / 600 * 1e6))
pstats.Stats(pr).sort_stats('cumulative').print_stats(18)
