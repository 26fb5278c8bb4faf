import sys, hashlib; sys.path.insert(0, '/root/repo')
import numpy as np, torch
from mddatasetbuilder_b200 import synth
from mddatasetbuilder_b200.engine import Engine
eng = Engine(0)
h = hashlib.sha256()
for natoms, dens, dense in [(20000, 0.02, False), (20000, 0.05, False), (8000, 0.11, True)]:
    s = synth.make_system(natoms, dens, seed=77, dense=dense); eng.set_elements(s.atomname)
    pos = synth.frame_positions(s, 2)
    counts, maxc, members = eng.compositions(pos, s.cell(), s.types, 5.0, members_cap=128)
    out = eng.descriptors(pos, s.cell(), s.types, 5.0, counts, maxc, members=members)
    eng.check()
    h.update(out['X'].cpu().numpy().tobytes())
print('X sha', h.hexdigest()[:16])
