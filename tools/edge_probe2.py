"""Developer tool: more odd inputs -- isolated targets, one-atom clusters, many element types, a small dense box
through DatasetBuilder."""
import os, sys, tempfile, shutil; sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch
from mddatasetbuilder_b200 import synth
from mddatasetbuilder_b200.engine import Engine
from oracle import oracle as O
fails = 0
def attempt(name, fn):
    global fails
    try:
        print("%-50s %s" % (name, fn()))
    except Exception as e:
        msg = str(e)
        if "CUDA error" in msg or "illegal" in msg: fails += 1
        print("%-50s raised: %s" % (name, msg[:140]))
eng = Engine(0, ("C", "H", "O"))
L = 40.0; cell = np.diag([L] * 3).reshape(9)
pos = np.array([[5.0, 5, 5], [25.0, 25, 25], [25.9, 25, 25], [5.0, 5, 5.5]])
types = np.array([0, 2, 1, 1], dtype=np.uint8)
def iso():
    counts, maxc = eng.compositions(pos[None], cell, types, 5.0); eng.check()
    out = eng.descriptors(pos[None], cell, types, 5.0, counts, maxc, want_eig=True, eigcap=4); eng.check()
    return counts.cpu().numpy().tolist(), np.round(out["eig"].cpu().numpy(), 4).tolist()
attempt("two-atom clusters far apart", iso)
def single():
    p1 = np.array([[1.0, 2, 3]]); t1 = np.array([2], dtype=np.uint8)
    counts, maxc = eng.compositions(p1[None], cell, t1, 5.0); eng.check()
    out = eng.descriptors(p1[None], cell, t1, 5.0, counts, maxc, want_eig=True, eigcap=2); eng.check()
    return counts.cpu().numpy().tolist(), out["eig"].cpu().numpy().tolist(), out["X"].cpu().numpy().tolist(), "expected diag", 0.5 * 8 ** 2.4
attempt("one-atom frame, one-atom cluster", single)
def many_types():
    e2 = Engine(0, ("H", "C", "N", "O", "F", "Si", "P", "S", "Cl"))
    return "accepted %d types" % len(e2.atomname)
attempt("nine element types", many_types)
def too_big_cluster():
    s = synth.make_system(3000, 0.11, seed=5, dense=True)
    eng.set_elements(s.atomname)
    counts, maxc = eng.compositions(s.pos0[None], s.cell(), s.types, 7.5, targets=np.arange(10, dtype=np.int32)); eng.check()
    out = eng.descriptors(s.pos0[None], s.cell(), s.types, 7.5, counts, maxc, targets=np.arange(10, dtype=np.int32)); eng.check()
    return "sizes", counts.cpu().numpy().sum(1).tolist()
attempt("clusters beyond 128 atoms", too_big_cluster)
eng.set_elements(("C", "H", "O"))
def small_box_builder():
    from mddatasetbuilder_b200 import DatasetBuilder
    s = synth.make_system(300, 0.10, seed=9, dense=True)
    d = tempfile.mkdtemp(prefix="mdb_small_"); cwd = os.getcwd(); os.chdir(d)
    try:
        synth.write_lammps_dump("dump.s", s, synth.frame_positions(s, 6, sigma=0.05))
        b = DatasetBuilder(atomname=list(s.atomname), dumpfilename="dump.s", dataset_name="s", n_clusters=20, stepinterval=1, seed=2)
        b.builddataset(writegjf=False)
        return "box %.1f A, %d types, %d structures" % (s.box[0], len(b.atombondtype), b._nstructure)
    finally:
        os.chdir(cwd); shutil.rmtree(d, ignore_errors=True)
attempt("DatasetBuilder on a 14 A box with 300 atoms", small_box_builder)
print("failures", fails)
