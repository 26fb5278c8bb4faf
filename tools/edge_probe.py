"""Developer tool: odd inputs through the public engine calls (tiny frames, atoms outside the cell, coincident
atoms, tiny cells, huge offsets, empty target lists), each compared with the oracle where it has an answer."""
import sys, traceback; sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch
from mddatasetbuilder_b200.engine import Engine
from oracle import oracle as O
from tests.util import ell_to_canonical
eng = Engine(0, ("C", "H", "O"))
Zt = np.array([6, 1, 8])
fails = 0
def check(name, pos, cell, types, periodic):
    global fails
    try:
        pos = np.asarray(pos, dtype=float); types = np.asarray(types, dtype=np.uint8)
        out = eng.analyze(pos[None], np.asarray(cell, dtype=float), types, periodic); eng.check()
        c3 = np.asarray(cell, dtype=float).reshape(3, 3) if np.size(cell) == 9 else np.diag(np.asarray(cell, dtype=float))
        bonds, _ = O.connect_the_dots(Zt[types].astype(np.int32), pos, c3, periodic, 1)
        rp, col = O.bonds_to_csr(len(types), bonds)
        crp, ccol = O.canonical_csr(rp, col)
        grp, gcol = ell_to_canonical(out["ell"].cpu().numpy()[0])
        _, _, lab = O.dps(rp, col)
        ok = np.array_equal(grp, crp) and np.array_equal(gcol, ccol) and np.array_equal(out["labels"].cpu().numpy()[0], lab)
        pr = eng.perceive(pos[None], np.asarray(cell, dtype=float) if periodic else None, types, out["ell"], periodic=periodic, labels=out["labels"]); eng.check()
        print("%-44s %s  bonds %d" % (name, "ok" if ok else "MISMATCH", len(bonds)))
        fails += not ok
    except Exception as e:
        fails += 1
        print("%-44s EXCEPTION %s" % (name, str(e)[:150]))
L = 20.0
cell = np.diag([L, L, L]).reshape(9)
rng = np.random.default_rng(0)
check("one atom, periodic", [[1, 2, 3]], cell, [0], True)
check("one atom, open", [[1, 2, 3]], cell, [2], False)
check("two coincident atoms", [[1, 2, 3], [1, 2, 3]], cell, [0, 0], True)
check("two atoms 0.39 A apart (below 0.4)", [[1, 2, 3], [1.39, 2, 3]], cell, [1, 1], True)
check("bond across the corner", [[0.1, 0.1, 0.1], [19.6, 19.7, 19.8]], cell, [2, 2], True)
p = rng.random((300, 3)) * L
check("atoms outside the cell (unwrapped +-3 L)", p + rng.integers(-3, 4, (300, 3)) * L, cell, rng.integers(0, 3, 300), True)
check("atoms exactly on the faces", np.array([[0, 0, 0], [L, L, L], [0, L, 0.96], [L, 0, 0]]) * 1.0, cell, [2, 2, 1, 1], True)
check("cell of 2.5 A (every axis a single cell)", rng.random((6, 3)) * 2.5, np.diag([2.5] * 3).reshape(9), rng.integers(0, 3, 6), True)
check("slab cell 40 x 40 x 3", rng.random((200, 3)) * [40, 40, 3], np.diag([40, 40, 3.0]).reshape(9), rng.integers(0, 3, 200), True)
check("open frame shifted by 1e6 A", p + 1e6, cell, rng.integers(0, 3, 300), False)
check("open frame, all atoms on a line", np.stack([np.arange(50) * 1.1, np.zeros(50), np.zeros(50)], 1), cell, [0] * 50, False)
check("strongly tilted cell", (rng.random((120, 3))) @ np.array([[12, 0, 0], [9, 12, 0], [7, -8, 12.0]]), np.array([[12, 0, 0], [9, 12, 0], [7, -8, 12.0]]).reshape(9), rng.integers(0, 3, 120), True)
# descriptor-side calls with nothing to do
try:
    types = rng.integers(0, 3, 300).astype(np.uint8)
    counts, maxc = eng.compositions(p[None], cell, types, 5.0, targets=np.zeros(0, dtype=np.int32))
    print("compositions with no targets: shape", tuple(counts.shape), "maxc", maxc.tolist() if hasattr(maxc, "tolist") else maxc)
except Exception as e:
    fails += 1; print("compositions with no targets: EXCEPTION", str(e)[:150])
try:
    counts, maxc = eng.compositions(p[None], cell, types, 12.0, targets=np.arange(5, dtype=np.int32))
    d = p[None, :5, :] - p[:, None, :]
    d -= L * np.round(d / L)
    want = (np.sqrt((d ** 2).sum(-1)) < 12.0).sum(0)
    got = counts.cpu().numpy().sum(1)
    fails += not np.array_equal(want, got)
    print("cutoff above half the cell: accepted, counts", got.tolist(), "minimum-image count", want.tolist())
except Exception as e:
    print("cutoff above half the cell: refused:", str(e)[:160])
try:
    counts, maxc = eng.compositions(p[None], cell, types, 10.0, targets=np.arange(5, dtype=np.int32))
    d = p[None, :5, :] - p[:, None, :]
    d -= L * np.round(d / L)
    want = (np.sqrt((d ** 2).sum(-1)) < 10.0).sum(0)
    fails += not np.array_equal(want, counts.cpu().numpy().sum(1))
    print("cutoff of exactly half the cell: counts", counts.cpu().numpy().sum(1).tolist(), "minimum-image count", want.tolist())
except Exception as e:
    fails += 1; print("cutoff of exactly half the cell: EXCEPTION", str(e)[:150])
def attempt(name, fn):
    global fails
    try:
        print("%-44s %s" % (name, fn()))
    except Exception as e:
        msg = str(e)
        if "CUDA error" in msg or "illegal" in msg:
            fails += 1
        print("%-44s raised: %s" % (name, msg[:110]))
empty = np.zeros(0, dtype=np.int32)
def desc_empty():
    counts, maxc = eng.compositions(p[None], cell, types, 5.0, targets=empty)
    out = eng.descriptors(p[None], cell, types, 5.0, counts, np.array([3, 3, 3], dtype=np.int32), targets=empty)
    return tuple(out["X"].shape)
attempt("descriptors with no targets", desc_empty)
attempt("sphere_members with no targets", lambda: [len(m) for m in eng.sphere_members(p[None], cell, types, 5.0, empty)])
X = torch.as_tensor(rng.random((50, 4)), device="cuda")
attempt("k-means assign, K = 1", lambda: eng.kmeans_assign(X, X[:1].contiguous())[0].cpu().numpy().max())
attempt("k-means assign, one row", lambda: eng.kmeans_assign(X[:1].contiguous(), X[:7].contiguous())[0].cpu().numpy().tolist())
attempt("k-means assign (tensor cores), K = 3", lambda: eng.kmeans_assign_tc(X, X[:3].contiguous())[0].cpu().numpy()[:5].tolist())
attempt("k-means++ with K = n", lambda: sorted(eng.kmeans_plusplus(X[:9].contiguous(), 9, rng.random((8, 4)), 2).cpu().numpy().tolist()))
Xc = X.clone(); Xc[:, 2] = 0.25
def mm():
    mn, mx = eng.minmax_fit(Xc); Y = Xc.clone(); eng.minmax_transform(Y, mn.cpu().numpy(), mx.cpu().numpy())
    return "constant column ->", Y[:, 2].unique().cpu().numpy().tolist()
attempt("min-max scaling of a constant column", mm)
print("failures", fails)
