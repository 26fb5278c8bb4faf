"""Developer tool: the three passes of the streaming job on ONE batch of 8 frames x 1M atoms (configs[2] density,
bond table), with a deliberately tiny clustering (K = 64, stop after 2 steps without improvement) so that the whole
process has ~2,000 launches -- the command the ncu launch list under profiles/ is taken from (the full bench job
has 300,000 launches, most of them mini-batch steps, which ncu cannot replay in any reasonable time).
Prints the live CUDA-event share of every kernel for comparison."""
import sys; sys.path.insert(0, '/root/repo')
import torch
from mddatasetbuilder_b200 import synth
from mddatasetbuilder_b200.engine import Engine
from mddatasetbuilder_b200.streaming import ResidentSource, StreamingSelector

F = 8
s = synth.make_system(1_000_000, 0.05, seed=synth.BASE_SEED + 3)
eng = Engine(0, s.atomname)
dev = eng.device
pos = synth.frame_positions_device(s, F, dev)
rowptr, col, bo = synth.bond_table_device(s, F, dev)
eng.profile(True)
sel = StreamingSelector(eng, s.types, n_clusters=64, seed=1, periodic=True, want_molecules=True, cache_gb=0.0, host_cache_gb=0.0,
                        kmeans_kwargs={"max_no_improvement": 2})
res = sel.run(ResidentSource(pos, s.cell(), F, rowptr=rowptr, col=col, bo=bo))
torch.cuda.synchronize()
prof = eng.profile_read()
tot = sum(ms for _, ms in prof.values())
print({k: round(v, 3) for k, v in res.timings.items()}, 'launches', eng.launches)
for k, (c, ms) in sorted(prof.items(), key=lambda kv: -kv[1][1])[:25]:
    print('%-28s %6d %9.2f ms %5.1f%%' % (k, c, ms, 100 * ms / tot))
