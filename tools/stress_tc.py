"""Developer tool: randomized sweep of the tensor-core assignment against the float64 kernel (whose labels
equal scikit-learn's in the tests): shapes, magnitudes, clustered / duplicated data."""
import sys, time; sys.path.insert(0, '/root/repo')
import numpy as np, torch
from mddatasetbuilder_b200.engine import Engine
eng = Engine(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 60
bad = 0; t0 = time.time(); worst = 0.0
for seed in range(n):
    rng = np.random.default_rng(seed)
    D = int(rng.integers(1, 113)); K = int(rng.choice([2, 7, 100, 129, 1000, 5000, 12000])); rows = int(rng.choice([1, 100, 4097, 30000]))
    kind = seed % 4
    if kind == 0:
        X = rng.random((rows, D)); C = rng.random((K, D))
    elif kind == 1:
        proto = rng.normal(0, 10.0 ** rng.integers(-2, 3), (max(K // 3, 1), D))
        X = proto[rng.integers(0, len(proto), rows)] + rng.normal(0, 1e-4, (rows, D))
        C = np.concatenate([proto, proto, proto])[:K] if K <= 3 * len(proto) else np.concatenate([proto] * (K // len(proto) + 1))[:K]
        C = C + rng.normal(0, 1e-6, C.shape) * (rng.random() < 0.5)
    elif kind == 2:
        X = rng.normal(500, 100, (rows, D)); C = X[rng.integers(0, rows, K)] + rng.normal(0, 1.0, (K, D))
    else:
        X = np.round(rng.random((rows, D)) * 4) / 4; C = np.round(rng.random((K, D)) * 4) / 4   # many exact ties
    Xd = torch.as_tensor(X, device='cuda'); Cd = torch.as_tensor(C, device='cuda')
    ref, _, _ = eng.kmeans_assign(Xd, Cd)
    got, nre = eng.kmeans_assign_tc(Xd, Cd)
    ok = bool(torch.equal(ref, got))
    worst = max(worst, nre / rows)
    if not ok:
        bad += 1
        print('MISMATCH seed', seed, 'D', D, 'K', K, 'rows', rows, 'kind', kind, int((ref != got).sum()))
print('cases', n, 'mismatches', bad, 'largest re-checked fraction %.3f' % worst, 'time %.0fs' % (time.time() - t0))
