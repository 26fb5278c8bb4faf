"""A/B of the descriptor kernels on one GPU: round-1 kernels (option desc_legacy = 1) against the current
ones, same targets; prints per-kernel CUDA-event times, ns per target and the largest row difference."""
import sys; sys.path.insert(0, '/root/repo')
import numpy as np, torch
from mddatasetbuilder_b200 import synth
from mddatasetbuilder_b200.engine import Engine

cases = [(200000, 0.05, False, 4)] if len(sys.argv) < 2 else [(int(sys.argv[1]), float(sys.argv[2]), False, int(sys.argv[3]))]
for natoms, dens, dense, F in cases:
    s = synth.make_system(natoms, dens, seed=synth.BASE_SEED + 3, dense=dense)
    eng = Engine(0, s.atomname)
    pos = synth.frame_positions_device(s, F, torch.device('cuda'))
    types = eng.to_device_types(s.types)
    counts, maxc, members = eng.compositions(pos, s.cell(), types, 5.0, members_cap=64)
    res = {}
    for mode in ((3,) if __import__('os').environ.get('DESC_AB_FINAL') else (1, 0, 3)):
        eng.set_option("desc_legacy", mode if mode < 3 else 0)
        eng.set_option("ql_pwk", 0 if mode < 3 else 1)
        for it in range(1 if __import__('os').environ.get('DESC_AB_ONCE') else 2):
            eng.profile(True)
            out = eng.descriptors(pos, s.cell(), types, 5.0, counts, maxc, members=members)
            torch.cuda.synchronize()
            pr = eng.profile_read(); eng.profile(False)
        eng.check()
        tri = sum(ms for k, (c, ms) in pr.items() if k.startswith('k_tridiag'))
        tql = sum(ms for k, (c, ms) in pr.items() if k.startswith('k_tql'))
        T = F * natoms
        print({1: 'legacy', 0: 'new   ', 3: 'new+pwk'}[mode], 'targets', T, 'tridiag %.2f ms (%.1f ns/target)  tql %.2f ms (%.1f ns/target)' % (tri, tri * 1e6 / T, tql, tql * 1e6 / T))
        print('   ', {k: round(ms, 2) for k, (c, ms) in sorted(pr.items()) if k.startswith('k_t')})
        res[mode] = out['X'].clone()
    n = out['n'].double()
    if len(res) < 3: continue
    print('max |X_pwk - X_new| = %.3e' % (res[3] - res[0]).abs().max().item())
    d = (res[0] - res[1]).abs().max().item()
    print('n mean %.1f max %d D %d | max |X_new - X_legacy| = %.3e (||M|| ~ %.1f)' % (n.mean().item(), n.max().item(), int(maxc.sum()), d, res[1].abs().max().item()))
    eng.close()
