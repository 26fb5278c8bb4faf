import sys; sys.path.insert(0, '/root/repo')
import numpy as np, torch
from mddatasetbuilder_b200 import synth
from mddatasetbuilder_b200.engine import Engine
from mddatasetbuilder_b200.kmeans import MiniBatchKMeansB200, MinMaxScalerB200
eng = Engine(0)
s = synth.make_system(3000, 0.05, seed=5)
eng.set_elements(s.atomname)
pos = synth.frame_positions(s, 4)
outs = []
for it in range(4):
    counts, maxc = eng.compositions(pos, s.cell(), s.types, 5.0)
    out = eng.descriptors(pos, s.cell(), s.types, 5.0, counts, maxc)
    outs.append(out['X'].clone())
for it in range(1, 4):
    d = (outs[it] - outs[0]).abs()
    print('descriptor run', it, 'bitwise equal', bool(torch.equal(outs[it], outs[0])), 'max abs diff', d.max().item())
X = outs[0]
labs = []
for it in range(4):
    km = MiniBatchKMeansB200(eng, n_clusters=50, init_size=150, n_init=3, seed=3).fit(X)
    labs.append(km.labels_.clone())
for it in range(1, 4):
    print('kmeans run', it, 'labels equal', bool(torch.equal(labs[it], labs[0])))
