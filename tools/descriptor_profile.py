import sys; sys.path.insert(0,'/root/repo')
import numpy as np, torch
from mddatasetbuilder_b200 import synth
from mddatasetbuilder_b200.engine import Engine
for natoms,dens,dense in [(100000,0.02,False),(200000,0.05,False),(100000,0.11,True)]:
    s=synth.make_system(natoms,dens,seed=synth.BASE_SEED+2,dense=dense)
    eng=Engine(0,s.atomname)
    pos=synth.frame_positions_device(s,4,torch.device('cuda'))
    types=eng.to_device_types(s.types)
    for it in range(2):
        eng.profile(True)
        counts,maxc,members=eng.compositions(pos,s.cell(),types,5.0,members_cap=128)
        out=eng.descriptors(pos,s.cell(),types,5.0,counts,maxc,members=members)
        torch.cuda.synchronize()
        pr=eng.profile_read(); eng.profile(False)
    n=out['n'].double()
    print(natoms,dens,'targets',4*natoms,'n mean %.1f max %d'%(n.mean().item(), n.max().item()),'D',int(maxc.sum()),{k:(c,round(ms,2)) for k,(c,ms) in pr.items()})
    eng.close()
