import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/oracle/shims')
import numpy as np, itertools
from mddatasetbuilder_b200 import synth
from mddatasetbuilder_b200.engine import Engine
from oracle import oracle as O
s=synth.make_system(900,0.06,seed=synth.BASE_SEED+55); L=float(s.box[0])
cell=np.array([[L,0,0],[3.0,L,0],[-2.0,4.0,L]])
pos=np.round((synth.frame_positions(s,1)[0]/L)@cell,6)[None]
eng=Engine(0,s.atomname)
tg=np.array([5],dtype=np.int32)
counts,maxc,members=eng.compositions(pos,cell.reshape(9),s.types,5.0,targets=tg,members_cap=96)
out=eng.descriptors(pos,cell.reshape(9),s.types,5.0,counts,maxc,targets=tg,members=members,want_eig=True,eigcap=96)
eng.check()
m=members.cpu().numpy()[0]; m=np.sort(m[m>=0]); n=len(m)
eig=out['eig'].cpu().numpy()[0,:n]
Z=np.array([O.ATOMIC_NUMBERS[s.atomname[t]] for t in s.types])
ortho=cell.T; frac=np.linalg.inv(ortho)
def mic(x):
    f=frac@x; f-=np.rint(f); b=ortho@f; best=b@b; o=b
    for h,k,l in itertools.product((-1,0,1),repeat=3):
        c=b+ortho@np.array([h,k,l],float)
        if c@c<best: best=c@c;o=c
    return o
P=pos[0]
vec=np.array([mic(P[j]-P[5]) for j in m])
def spec(pairfun):
    M=np.zeros((n,n))
    for i in range(n):
        M[i,i]=Z[m[i]]**2.4/2
        for j in range(i):
            d=pairfun(vec[j]-vec[i]); r=np.linalg.norm(d); M[i,j]=M[j,i]=Z[m[i]]*Z[m[j]]/r
    return np.linalg.eigvalsh(M)
print('n',n,'gpu',eig[:3],eig[-3:])
print('true mic ',spec(mic)[[0,1,2,-3,-2,-1]])
print('no pairmic',spec(lambda d:d)[[0,1,2,-3,-2,-1]])
Ld=np.array([L,L,L])
print('ortho wrap',spec(lambda d:d-Ld*np.rint(d/Ld))[[0,1,2,-3,-2,-1]])
vec2=np.array([(P[j]-P[5])-Ld*np.rint((P[j]-P[5])/Ld) for j in m])
vec_save=vec; vec=vec2
print('ortho gather + true pair',spec(mic)[[0,1,2,-3,-2,-1]])
print('ortho gather + no pair',spec(lambda d:d)[[0,1,2,-3,-2,-1]])
print('trace gpu', eig.sum(), 'true', sum(Z[m]**2.4/2))
raw=np.array([P[j]-P[5] for j in m]); vec=raw
print('raw gather + true pair', spec(mic)[[0,1,2,-3,-2,-1]])
print('raw gather + no pair', spec(lambda d:d)[[0,1,2,-3,-2,-1]])
for mode in (1,2):
    eng.set_option('desc_legacy',mode)
    o2=eng.descriptors(pos,cell.reshape(9),s.types,5.0,counts,maxc,targets=tg,members=members,want_eig=True,eigcap=96)
    eng.check(); print('mode',mode,o2['eig'].cpu().numpy()[0,:3], o2['eig'].cpu().numpy()[0,n-3:n])
# an orthorhombic control with the same code path
cell0=np.diag([L,L,L]); pos0=np.round(synth.frame_positions(s,1),6)
c0,m0,mem0=eng.compositions(pos0,cell0.reshape(9),s.types,5.0,targets=tg,members_cap=96)
o3=eng.descriptors(pos0,cell0.reshape(9),s.types,5.0,c0,m0,targets=tg,members=mem0,want_eig=True,eigcap=96); eng.check()
print('ortho control', o3['eig'].cpu().numpy()[0,:3])
