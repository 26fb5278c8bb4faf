"""Developer tool: where the time of the device dump parser goes (steady state, staging buffer reused)."""
import os, sys, time, tempfile, shutil; sys.path.insert(0, '/root/repo')
import numpy as np, torch
from mddatasetbuilder_b200 import synth
from mddatasetbuilder_b200.engine import Engine
from mddatasetbuilder_b200.reader import KIND_DUMP, TextReader
natoms, nf, B = 100000, 32, 8
s = synth.make_system(natoms, 0.02, seed=3)
eng = Engine(0, s.atomname)
d = tempfile.mkdtemp(prefix='mdb_tp_')
fn = os.path.join(d, 'dump.big')
synth.write_lammps_dump(fn, s, synth.frame_positions(s, nf))
size = os.path.getsize(fn)
r = TextReader(KIND_DUMP, fn, nthreads=0)
for it in range(3):
    r.rewind(); eng.profile(it == 2)
    torch.cuda.synchronize(); t0 = time.perf_counter(); got = 0
    while True:
        pos, cell = r.read_device(B, eng, prefetch=True)
        if pos.shape[0] == 0: break
        got += pos.shape[0]
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print('device parser pass %d: %.1f ms  %.0f MB/s  %.3e atom-frames/s' % (it, dt * 1e3, size / 1e6 / dt, got * natoms / dt))
print({k: (c, round(ms, 3)) for k, (c, ms) in eng.profile_read().items()})
buf = np.empty((B, natoms, 3))
for it in range(2):
    r.rewind(); t0 = time.perf_counter(); got = 0
    while True:
        pos, cell = r.read(B, pinned=buf)
        if len(pos) == 0: break
        got += len(pos)
    dt = time.perf_counter() - t0
    print('host parser pass %d: %.1f ms  %.0f MB/s  %.3e atom-frames/s' % (it, dt * 1e3, size / 1e6 / dt, got * natoms / dt))
shutil.rmtree(d, ignore_errors=True)
