"""Developer tool: wall time of a DatasetBuilder run on a synthetic trajectory (dump-only mode), per step,
with the per-kernel GPU time beside it."""
import os, sys, time, tempfile, shutil, logging; sys.path.insert(0, '/root/repo')
import numpy as np, torch
from mddatasetbuilder_b200 import synth, DatasetBuilder
natoms, nf = int(sys.argv[1]) if len(sys.argv) > 1 else 100000, int(sys.argv[2]) if len(sys.argv) > 2 else 40
mode = sys.argv[3] if len(sys.argv) > 3 else 'dump'
s = synth.make_system(natoms, 0.02, seed=11)
d = tempfile.mkdtemp(prefix='mdb_build_'); os.chdir(d)
t0 = time.time(); synth.write_lammps_dump('dump.syn', s, synth.frame_positions(s, nf))
if mode == 'bonds': synth.write_lammps_bonds('bonds.syn', s, nf)
print('wrote inputs in %.1f s (%.0f MB)' % (time.time() - t0, os.path.getsize('dump.syn') / 1e6))
logging.basicConfig(level=logging.INFO)
b = DatasetBuilder(atomname=list(s.atomname), dumpfilename='dump.syn', bondfilename='bonds.syn' if mode == 'bonds' else None,
                   dataset_name='syn', n_clusters=10000, stepinterval=1, seed=1)
eng = b.engine; eng.profile(True)
import cProfile, pstats, io
pr_ = cProfile.Profile() if os.environ.get('MDB_CPROFILE') else None
t0 = time.time()
if pr_: pr_.enable()
b.builddataset(writegjf=bool(os.environ.get('MDB_WRITEGJF'))); torch.cuda.synchronize()
if pr_: pr_.disable()
dt = time.time() - t0
if pr_:
    st_ = io.StringIO(); pstats.Stats(pr_, stream=st_).sort_stats('cumulative').print_stats(32); print(st_.getvalue()[-6500:])
print('builddataset: %.2f s for %d atom-frames (%.3e atom-frames/s), %d structures' % (dt, natoms * nf, natoms * nf / dt, b._nstructure))
pr = eng.profile_read(); tot = sum(ms for _, ms in pr.values())
print('GPU kernel time %.2f s' % (tot / 1e3))
for k, (c, ms) in sorted(pr.items(), key=lambda kv: -kv[1][1])[:12]: print('  %-24s %7d launches %9.1f ms' % (k, c, ms))
os.chdir('/'); shutil.rmtree(d, ignore_errors=True)
