"""Developer tool: randomized parity sweep of the bond path against the oracle (many seeds, densities,
reactive fractions, noise levels, cell shapes).  Not part of the test suite (minutes of oracle time)."""
import sys, time; sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch
from mddatasetbuilder_b200 import synth
from mddatasetbuilder_b200.engine import Engine
from oracle import oracle as O
from util import ell_to_canonical, oracle_frame
eng = Engine(0, ("C", "H", "O"))
nseeds = int(sys.argv[1]) if len(sys.argv) > 1 else 30
bad = 0; ncase = 0; wide = 0; ndel_tot = 0; t0 = time.time()
for seed in range(nseeds):
    rng = np.random.default_rng(1000 + seed)
    natoms = int(rng.choice([600, 1500, 3000, 5000]))
    density = float(rng.choice([0.02, 0.05, 0.08, 0.11]))
    dense = density >= 0.08
    rf = float(rng.choice([0.01, 0.05, 0.2]))
    s = synth.make_system(natoms, density, seed=5000 + seed, reactive_fraction=rf, dense=dense)
    noise = float(rng.choice([0.02, 0.1, 0.3]))
    pos = s.pos0 + rng.normal(0, noise, s.pos0.shape)
    kind = seed % 3
    if kind == 0:
        cell, periodic, p = s.cell(), True, pos
    elif kind == 1:
        cell, periodic, p = s.cell(), False, pos
    else:
        L = s.box[0]
        cell = np.array([[L, 0, 0], [rng.uniform(-0.4, 0.4) * L, L, 0], [rng.uniform(-0.3, 0.3) * L, rng.uniform(-0.3, 0.3) * L, L]])
        periodic, p = True, (pos / L) @ cell
    try:
        out = eng.analyze(p[None], cell, s.types, periodic)
        st = eng.check()
    except RuntimeError as e:   # more than 8 raw neighbours somewhere: the documented remedy
        assert "max_bonds" in str(e), e
        eng.set_option("max_bonds", 16)
        out = eng.analyze(p[None], cell, s.types, periodic)
        st = eng.check()
        eng.set_option("max_bonds", 8)
        wide += 1
    rp, col, ndel = oracle_frame(s, p, periodic=periodic, mode=0, cell=cell)
    ndel_tot += ndel
    grp, gcol = ell_to_canonical(out["ell"].cpu().numpy()[0])
    crp, ccol = O.canonical_csr(rp, col)
    _, _, lab = O.dps(rp, col)
    ok = np.array_equal(grp, crp) and np.array_equal(gcol, ccol) and np.array_equal(out["labels"].cpu().numpy()[0], lab)
    ncase += 1
    if not ok:
        bad += 1
        print('MISMATCH seed', seed, natoms, density, rf, noise, kind)
print('cases', ncase, 'mismatches', bad, 'cases that needed max_bonds=16:', wide, 'bonds deleted by the cleanup', ndel_tot, 'time %.0fs' % (time.time() - t0))
