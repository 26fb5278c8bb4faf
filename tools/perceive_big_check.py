"""Developer tool: one bench-sized frame set (100k atoms) through the perception, against the restatement."""
import sys, time; sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
from mddatasetbuilder_b200 import synth
from mddatasetbuilder_b200.engine import Engine
from test_perceive_gpu import _oracle_orders, ZNUM
eng = Engine(0, ("C", "H", "O"))
natoms = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
for density, dense in ((0.02, False), (0.10, True)):
    s = synth.make_system(natoms, density, seed=synth.BASE_SEED + 2, dense=dense)
    frames = synth.frame_positions(s, 2, sigma=0.08)
    cell = np.diag(s.box).reshape(9)
    out = eng.analyze(frames, cell, s.types); eng.check()
    got = eng.perceive(frames, cell, s.types, out["ell"], labels=out["labels"]); rings = eng.perceive_stats()["rings"]; eng.check()
    Z = [ZNUM[s.atomname[t]] for t in s.types]
    t0 = time.time(); bad = 0; nm = 0; nr = 0
    for f in range(2):
        want, info = _oracle_orders(out["ell"].cpu().numpy()[f], frames[f], Z, cell, True)
        bad += int((want != got["orders"].cpu().numpy()[f]).sum()); nm += int((want > 1).sum()) // 2; nr += info["rings"]
    print("density %.2f: %d atoms x 2 frames, multiple bonds %d, rings %d (device %d), slots that differ: %d, oracle %.0f s"
          % (density, natoms, nm, nr, rings, bad, time.time() - t0))
