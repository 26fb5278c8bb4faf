"""Developer tool: k-means++ seeding on the 30,000-row pool DatasetBuilder seeds from (3 seedings at once)."""
import sys, time; sys.path.insert(0, '/root/repo')
import numpy as np, torch
from mddatasetbuilder_b200.engine import Engine
eng = Engine(0, ("C", "H", "O"))
n, D, K, B = int(sys.argv[1]) if len(sys.argv) > 1 else 30000, 35, 10000, 3
T = 2 + int(np.log(K))
rng = np.random.default_rng(0)
X = torch.as_tensor(rng.random((200000, D)), device="cuda")
rowidx = torch.as_tensor(rng.integers(0, 200000, (B, n)).astype(np.int32), device="cuda")
rand = rng.random((B, K - 1, T)); first = rng.integers(0, n, B)
for it in range(3):
    torch.cuda.synchronize(); t0 = time.time()
    idx = eng.kmeans_plusplus(X, K, rand, first, rowidx=rowidx)
    torch.cuda.synchronize(); dt = time.time() - t0
print("n=%d: %.1f ms per %d x %d steps -> %.2f us per step" % (n, dt * 1e3, B, K, dt / K * 1e6))
