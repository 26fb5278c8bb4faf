import sys; sys.path.insert(0, '/root/repo')
import numpy as np
from mddatasetbuilder_b200.engine import Engine
from oracle import oracle as O
eng = Engine(0, ("C", "H", "O"))
rng = np.random.default_rng(0)
for L, N in ((20.0, 300), (14.0, 280), (20.0, 2500), (30.0, 300)):
    p = rng.random((N, 3)) * L
    types = rng.integers(0, 3, N).astype(np.uint8)
    cell = np.diag([L, L, L]).reshape(9)
    for cutoff in (3.0, 5.0, 6.6, 6.7, 8.0, 9.9, 10.0):
        if 2 * cutoff > L: continue
        counts, _ = eng.compositions(p[None], cell, types, cutoff, targets=np.arange(8, dtype=np.int32))
        eng.check()
        got = counts.cpu().numpy().sum(1)
        want = np.array([len(O.sphere(p, np.array([L, L, L]), True, a, cutoff)) for a in range(8)])
        print(L, N, cutoff, "ok" if np.array_equal(got, want) else "DIFF", got.tolist(), want.tolist())
