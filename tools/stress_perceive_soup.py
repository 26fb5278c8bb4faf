"""Developer tool: perception parity on random atom soups (uniform random C/H/O positions at liquid-like
density): after the valence cleanup these are tangled networks full of rings, sp / sp2 chains and odd
valences -- an adversarial input for the ordered pass-6 turns."""
import sys, time; sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
from mddatasetbuilder_b200.engine import Engine
from test_perceive_gpu import _oracle_orders, ZNUM
names = ("C", "H", "O")
eng = Engine(0, names)
nseeds = int(sys.argv[1]) if len(sys.argv) > 1 else 20
bad = 0; nm = 0; nt = 0; nr = 0; wide = 0; t0 = time.time()
for seed in range(nseeds):
    rng = np.random.default_rng(9000 + seed)
    n = int(rng.choice([500, 1500, 3000]))
    density = float(rng.choice([0.05, 0.08, 0.11]))
    L = (n / density) ** (1 / 3)
    frac_h = float(rng.choice([0.0, 0.3, 0.6]))
    types = rng.choice(3, n, p=[(1 - frac_h) * 0.7, frac_h, (1 - frac_h) * 0.3]).astype(np.uint8)
    pos = rng.random((n, 3)) * L
    # push apart pairs closer than 0.9 A (ConnectTheDots ignores d < 0.4 A; overlapping atoms are not the point)
    for _ in range(3):
        d = pos[:, None, :] - pos[None, :, :]
        d -= L * np.round(d / L)
        r = np.sqrt((d ** 2).sum(-1)) + np.eye(n) * 10
        i, j = np.nonzero(r < 0.9)
        for a, b in zip(i, j):
            if a < b:
                pos[b] += d[b, a] / r[b, a] * (1.0 - r[b, a])
        pos %= L
    cell = np.diag([L, L, L]).reshape(9)
    periodic = seed % 4 != 3
    mb = 8
    try:
        out = eng.analyze(pos[None], cell, types, periodic); eng.check()
    except RuntimeError as e:
        assert "max_bonds" in str(e), e
        mb = 16; wide += 1
        eng.set_option("max_bonds", 16)
        out = eng.analyze(pos[None], cell, types, periodic); eng.check()
    got = eng.perceive(pos[None], cell if periodic else None, types, out["ell"], periodic=periodic, labels=out["labels"])
    rings = eng.perceive_stats()["rings"]; eng.check()
    if mb == 16: eng.set_option("max_bonds", 8)
    Z = [ZNUM[names[t]] for t in types]
    want, info = _oracle_orders(out["ell"].cpu().numpy()[0], pos, Z, cell if periodic else None, periodic)
    o = got["orders"].cpu().numpy()[0]
    nm += int((want == 2).sum()) // 2; nt += int((want == 3).sum()) // 2; nr += info["rings"]
    if not np.array_equal(want, o) or rings != info["rings"]:
        bad += 1
        d = np.argwhere(want != o)
        print("MISMATCH seed", seed, n, density, frac_h, "slots", len(d), d[:4].tolist(), "rings", rings, info["rings"])
print("soups", nseeds, "mismatches", bad, "double bonds", nm, "triple bonds", nt, "rings", nr, "max_bonds=16 cases", wide, "time %.0fs" % (time.time() - t0))
