"""Developer tool: randomized parity sweep of the bond-order perception (csrc/perceive.cu) against the
restatement (oracle/perceive.py): many seeds, densities, noise levels, cell shapes, both row widths.
Dense noisy systems give tangled networks (rings, chains of sp2 carbons) that exercise the ordered
pass-6 turns.  Not part of the test suite (minutes of oracle time)."""
import sys, time; sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
from mddatasetbuilder_b200 import synth
from mddatasetbuilder_b200.engine import Engine
from test_perceive_gpu import _oracle_orders, ZNUM
eng = Engine(0, ("C", "H", "O"))
nseeds = int(sys.argv[1]) if len(sys.argv) > 1 else 30
bad = 0; ncase = 0; wide = 0; nmult = 0; nring = 0; ntrip = 0; t0 = time.time()
for seed in range(nseeds):
    rng = np.random.default_rng(3000 + seed)
    natoms = int(rng.choice([400, 1000, 2000]))
    density = float(rng.choice([0.03, 0.06, 0.09, 0.12]))
    dense = density >= 0.09
    s = synth.make_system(natoms, density, seed=7000 + seed, reactive_fraction=float(rng.choice([0.0, 0.05, 0.2])), dense=dense)
    noise = float(rng.choice([0.02, 0.08, 0.2, 0.35]))
    pos = s.pos0 + rng.normal(0, noise, s.pos0.shape)
    if seed % 5 == 4:   # a soup of carbons and oxygens squeezed together: sp / sp2 chains and rings
        sel = s.types != 1
        pos = pos.copy(); pos[sel] *= 0.8
    kind = seed % 3
    if kind == 0:
        cell, periodic, p = s.cell(), True, pos
    elif kind == 1:
        cell, periodic, p = s.cell(), False, pos
    else:
        L = s.box[0]
        cell = np.array([[L, 0, 0], [rng.uniform(-0.4, 0.4) * L, L, 0], [rng.uniform(-0.3, 0.3) * L, rng.uniform(-0.3, 0.3) * L, L]])
        periodic, p = True, (pos / L) @ cell
    mb = 8
    try:
        out = eng.analyze(p[None], cell, s.types, periodic)
        eng.check()
    except RuntimeError as e:
        assert "max_bonds" in str(e), e
        mb = 16
        eng.set_option("max_bonds", 16)
        out = eng.analyze(p[None], cell, s.types, periodic)
        eng.check()
        wide += 1
    got = eng.perceive(p[None], cell if periodic else None, s.types, out["ell"], periodic=periodic, labels=out["labels"])
    ring = eng.perceive_stats()["rings"]
    eng.check()
    if mb == 16:
        eng.set_option("max_bonds", 8)
    Z = [ZNUM[s.atomname[t]] for t in s.types]
    want, info = _oracle_orders(out["ell"].cpu().numpy()[0], p, Z, np.asarray(cell, dtype=float).reshape(-1) if periodic else None, periodic)
    orders = got["orders"].cpu().numpy()[0]
    ncase += 1
    nmult += int((want > 1).sum()) // 2
    ntrip += int((want == 3).sum()) // 2
    nring += info["rings"]
    if not np.array_equal(want, orders) or ring != info["rings"]:
        bad += 1
        d = np.argwhere(want != orders)
        print('MISMATCH seed', seed, natoms, density, noise, kind, 'slots', len(d), 'first', d[:3].tolist(), 'rings', ring, info["rings"])
print('cases', ncase, 'mismatches', bad, 'max_bonds=16 cases', wide, 'multiple bonds', nmult, 'triples', ntrip, 'ring molecules', nring,
      'time %.0fs' % (time.time() - t0))
