import sys; sys.path.insert(0,'/root/repo')
import numpy as np, torch
from mddatasetbuilder_b200 import synth
from mddatasetbuilder_b200.engine import Engine
s=synth.make_system(200000,0.05,seed=synth.BASE_SEED+2)
eng=Engine(0,s.atomname)
pos=synth.frame_positions_device(s,1,torch.device('cuda'))
types=eng.to_device_types(s.types)
for it in range(2):
    counts,maxc=eng.compositions(pos,s.cell(),types,5.0)
    out=eng.descriptors(pos,s.cell(),types,5.0,counts,maxc)
    torch.cuda.synchronize()
print('ok', out['n'].double().mean().item())
