"""Developer tool: per-kernel time of the bond-order perception on bench-like frames."""
import sys; sys.path.insert(0, '/root/repo')
import numpy as np, torch
from mddatasetbuilder_b200 import synth
from mddatasetbuilder_b200.engine import Engine
eng = Engine(0, ("C", "H", "O"))
natoms, frames = 100000, int(sys.argv[1]) if len(sys.argv) > 1 else 40
s = synth.make_system(natoms, 0.02, seed=synth.BASE_SEED + 2)
pos = synth.frame_positions_device(s, frames, torch.device("cuda:0"), sigma=0.05)
cell = np.diag(s.box).reshape(9)
out = eng.analyze(pos, cell, s.types)
eng.check()
for it in range(3):
    eng.profile(it == 2)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    r = eng.perceive(pos, cell, s.types, out["ell"], labels=out["labels"])
    ev1.record(); torch.cuda.synchronize()
print("perceive %.3f ms for %d atom-frames -> %.2e atom-frames/s" % (ev0.elapsed_time(ev1), natoms * frames, natoms * frames / ev0.elapsed_time(ev1) * 1e3))
for k, v in eng.profile_read().items():
    print(k, v)
eng.profile(False)
o = r["orders"]
print("multiple bonds", int((o > 1).sum()) // 2, "pending counters", None)
eng.set_option("perceive_bond_orders", 1)
for it in range(3):
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(); eng.analyze(pos, cell, s.types); ev1.record(); torch.cuda.synchronize()
print("analyze with perceived keys %.3f ms" % ev0.elapsed_time(ev1))
eng.set_option("perceive_bond_orders", 0)
for it in range(3):
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(); eng.analyze(pos, cell, s.types); ev1.record(); torch.cuda.synchronize()
print("analyze with all-single keys %.3f ms" % ev0.elapsed_time(ev1))
