import os, sys, time, tempfile, shutil, ctypes as C; sys.path.insert(0, '/root/repo')
import numpy as np, torch
from mddatasetbuilder_b200 import synth, _lib
from mddatasetbuilder_b200.reader import KIND_DUMP, TextReader
natoms, nf, B = 100000, 32, 8
s = synth.make_system(natoms, 0.02, seed=3)
d = tempfile.mkdtemp(prefix='mdb_tp_'); fn = os.path.join(d, 'dump.big')
synth.write_lammps_dump(fn, s, synth.frame_positions(s, nf)); size = os.path.getsize(fn)
r = TextReader(KIND_DUMP, fn, nthreads=0)
stage = torch.empty(int(size / nf * 1.3) * B, dtype=torch.uint8, pin_memory=True)
fb = np.empty(B, np.int64); ab = np.empty(B, np.int64); fe = np.empty(B, np.int64); cell = np.empty((B, 9)); nread = C.c_int(0)
dtext = torch.empty(stage.numel(), dtype=torch.uint8, device='cuda')
for it in range(3):
    r.rewind(); t_raw = 0; t_h2d = 0
    while True:
        t0 = time.perf_counter()
        rc = r.L.mdb_reader_next_raw(r.h, 1, B, C.c_void_p(stage.data_ptr()), C.c_size_t(stage.numel()), C.c_void_p(fb.ctypes.data), C.c_void_p(ab.ctypes.data), C.c_void_p(fe.ctypes.data), C.c_void_p(cell.ctypes.data), C.byref(nread))
        t1 = time.perf_counter(); t_raw += t1 - t0
        if nread.value == 0: break
        tot = int(fe[nread.value - 1])
        dtext[:tot].copy_(stage[:tot], non_blocking=True); torch.cuda.synchronize()
        t_h2d += time.perf_counter() - t1
    print('pass', it, 'next_raw %.1f ms  h2d %.1f ms' % (t_raw * 1e3, t_h2d * 1e3))
shutil.rmtree(d, ignore_errors=True)
