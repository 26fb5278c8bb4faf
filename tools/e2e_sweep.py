"""Developer tool: host-buffer path timing against the chunk size (rows variant, then the bond-list variant;
MDB_SWEEP_NOOUT=1 drops the labels and key codes too, i.e. input copy + kernels + the bond list only)."""
import os
import sys, time; sys.path.insert(0, '/root/repo')
import numpy as np, torch
from mddatasetbuilder_b200 import synth
from mddatasetbuilder_b200.engine import Engine
natoms, frames = 100000, 1000
s = synth.make_system(natoms, 0.02, seed=synth.BASE_SEED + 2)
eng = Engine(0, s.atomname)
pos = synth.frame_positions_device(s, frames, torch.device('cuda'))
h_pos = torch.empty((frames, natoms, 3), dtype=torch.float64, pin_memory=True); h_pos.copy_(pos); del pos
h_out = {"ell": torch.empty((frames, natoms, 4), dtype=torch.int32, pin_memory=True),
         "keys": torch.empty((frames, natoms), dtype=torch.int64, pin_memory=True),
         "labels": torch.empty((frames, natoms), dtype=torch.int32, pin_memory=True)}
for fpl in [int(a) for a in sys.argv[1:]] or [160, 80, 40, 20]:
    eng.set_option("frames_per_launch", fpl)
    eng.analyze_host(h_pos, s.cell(), s.types, True, out=h_out, ell_width=4)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3): eng.analyze_host(h_pos, s.cell(), s.types, True, out=h_out, ell_width=4)
    torch.cuda.synchronize(); ms = (time.perf_counter() - t0) / 3 * 1e3
    print('frames_per_launch %4d: %.1f ms/step  %.3e atom-frames/s  (%.1f GB/s in, %.1f GB/s out)' % (
        fpl, ms, natoms * frames / ms * 1e3, natoms * frames * 24 / ms / 1e6, natoms * frames * 28 / ms / 1e6))

h_e = {"edges": torch.empty((frames * natoms, 2), dtype=torch.int32, pin_memory=True), "edge_ptr": torch.empty(frames + 1, dtype=torch.int64, pin_memory=True),
       "keycode": torch.empty((frames, natoms), dtype=torch.uint8, pin_memory=True), "keytable": torch.empty(256, dtype=torch.int64, pin_memory=True),
       "labels": h_out["labels"]}
noout = bool(os.environ.get("MDB_SWEEP_NOOUT"))
kw = dict(want_keys=not noout, want_labels=not noout)
for fpl in [int(a) for a in sys.argv[1:]] or [160, 80, 40, 20]:
    eng.set_option("frames_per_launch", fpl)
    eng.analyze_host(h_pos, s.cell(), s.types, True, out=h_e, edges=True, **kw)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3): eng.analyze_host(h_pos, s.cell(), s.types, True, out=h_e, edges=True, **kw)
    torch.cuda.synchronize(); ms = (time.perf_counter() - t0) / 3 * 1e3
    print('bond list, frames_per_launch %4d: %.1f ms/step  %.3e atom-frames/s  (%.1f GB/s in)' % (fpl, ms, natoms * frames / ms * 1e3, natoms * frames * 24 / ms / 1e6))
