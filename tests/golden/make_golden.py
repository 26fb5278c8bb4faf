"""Generate the committed golden fixtures by running the reference's OWN, unmodified Python.

Run in the build container only (needs /root/reference; `make -C oracle` first):

    python tests/golden/make_golden.py

What runs: /root/reference/mddatasetbuilder/{detect,datasetbuilder,utils}.py and its compiled
Cython `dps` (oracle/_ref), imported through oracle/refrun.py.  The four third-party packages the
reference needs and this image lacks (ase, openbabel, lz4, coloredlogs) are the stand-ins under
oracle/shims, whose MIC / ConnectTheDots arithmetic is a restatement -- so the
fixtures pin the reference's host logic (parsing, keys, row assembly, molecule completion, file
naming, text formats) exactly, and its third-party arithmetic as restated.

Outputs (all small, committed):
  tests/golden/traj/dump.small, bonds.small      LAMMPS-format inputs (synthetic, seeded)
  tests/golden/golden.json                       per-frame method outputs + dataset manifest
  tests/golden/dataset_sample/                   a few full .xyz / .gjf outputs
"""
import hashlib
import json
import os
import pickle
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from mddatasetbuilder_b200 import synth  # noqa: E402
from oracle import refrun  # noqa: E402


def sha(path):
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def split_inputs(dump, bonds, dump_lines, bond_lines, outdir, first=3):
    """The 6-frame inputs as two files each (3 + 3 frames): ([dump parts], [bond parts])."""
    out = []
    for path, n, tag in ((dump, dump_lines, "dump"), (bonds, bond_lines, "bonds")):
        with open(path) as f:
            lines = f.readlines()
        names = [os.path.join(outdir, f"{tag}.part{k}") for k in (1, 2)]
        with open(names[0], "w") as f:
            f.writelines(lines[: first * n])
        with open(names[1], "w") as f:
            f.writelines(lines[first * n:])
        out.append(names)
    return out


def main():
    ref = refrun.load()
    from mddatasetbuilder.datasetbuilder import DatasetBuilder
    from mddatasetbuilder.detect import DetectBond, DetectDump

    traj = os.path.join(HERE, "traj")
    os.makedirs(traj, exist_ok=True)
    s = synth.make_system(96, 0.045, seed=synth.BASE_SEED + 101, reactive_fraction=0.08)
    frames = synth.frame_positions(s, 6, sigma=0.04)
    dump = os.path.join(traj, "dump.small")
    bonds = os.path.join(traj, "bonds.small")
    synth.write_lammps_dump(dump, s, frames, shuffle_seed=5)
    # per-frame bond orders so that rounding / keys vary between frames
    rng = np.random.default_rng(9)
    bo_frames = [np.round(np.clip(s.bo + rng.uniform(-0.25, 0.25, s.bo.shape), 0.05, None), 3) for _ in range(6)]
    # keep both directions of a bond equal
    for bf in bo_frames:
        for i in range(s.natoms):
            for p in range(s.rowptr[i], s.rowptr[i + 1]):
                j = s.col[p]
                if j > i:
                    q = s.rowptr[j] + list(s.col[s.rowptr[j]:s.rowptr[j + 1]]).index(i)
                    bf[q] = bf[p]
    synth.write_lammps_bonds(bonds, s, 6, bo_frames=bo_frames)

    atomname = np.array(s.atomname)
    g = {"atomname": list(s.atomname), "natoms": s.natoms, "box": s.box.tolist(), "nframes": 6, "frames": []}
    dd = DetectDump(filename=dump, atomname=atomname, pbc=True)
    db = DetectBond(filename=bonds, atomname=atomname, pbc=True)
    g["dump_steplinenum"] = dd.steplinenum
    g["bond_steplinenum"] = db.steplinenum
    g["atomtype"] = dd.atomtype.tolist()

    def frames_of(path, n):
        with open(path) as f:
            lines = f.readlines()
        return [tuple(lines[k * n:(k + 1) * n]) for k in range(len(lines) // n)]

    dframes = frames_of(dump, dd.steplinenum)
    bframes = frames_of(bonds, db.steplinenum)

    def decode(d):
        return sorted(([str(pickle.loads(k)[0]), [int(x) for x in pickle.loads(k)[1]], [int(i) for i in v]]
                       for k, v in d.items()), key=lambda t: t[2][0])

    for k in range(6):
        fr = {}
        atoms, ids = dd.readcrd(dframes[k])
        fr["cell"] = np.asarray(atoms.cell).tolist()
        fr["pos_sha"] = hashlib.sha256(np.ascontiguousarray(atoms.positions).tobytes()).hexdigest()
        d, step = dd.readatombondtype(((k, dframes[k]), None))
        fr["dump_keys"] = decode(d)
        d, step = db.readatombondtype(((k, bframes[k]), None))
        fr["bond_keys"] = decode(d)
        mols, _ = dd.readmolecule(dframes[k])
        fr["dump_molecules"] = [[int(a) for a in m] for m in mols]
        fr["dump_adjacency"] = [[int(a) for a in r] for r in DetectDump._crd2bond(atoms, False)]
        mols, _ = db.readmolecule(bframes[k])
        fr["bond_molecules"] = [[int(a) for a in m] for m in mols]
        g["frames"].append(fr)

    # descriptors of one bond type in frame 1 through the reference's _writestepmatrix
    cwd = os.getcwd()
    tmp = tempfile.mkdtemp()
    os.chdir(tmp)
    try:
        b = DatasetBuilder(atomname=list(s.atomname), bondfilename=bonds, dumpfilename=dump, dataset_name="gold",
                           cutoff=5.0, stepinterval=2, n_clusters=10000, nproc=1)
        b.dstep = {1: [int(i) for i in range(1, 25)]}
        res = b._writestepmatrix((1, dframes[1]))
        g["stepmatrix_frame1"] = [{"atom": int(sa[1]), "eig": [float(x) for x in e],
                                   "counter": {str(kk): int(vv) for kk, vv in c.items()}} for sa, e, c in res]
        # end-to-end, bonds + dump, every row taken (n_atoms <= n_clusters): deterministic outputs
        b = DatasetBuilder(atomname=list(s.atomname), bondfilename=bonds, dumpfilename=dump, dataset_name="gold",
                           cutoff=5.0, stepinterval=2, n_clusters=10000, nproc=1,
                           qmkeywords="%nproc=4\n#force mn15/6-31g** Geom=PrintInputOrient")
        b.builddataset()
        manifest = {}
        for base in (b.dataset_dir, b.gjfdir):
            for dirpath, _, files in os.walk(base):
                for fn in files:
                    p = os.path.join(dirpath, fn)
                    manifest[os.path.relpath(p, tmp)] = sha(p)
        g["e2e_bonds"] = {"atombondtype": list(b.atombondtype), "nstructure": int(b._nstructure), "manifest": manifest}
        sample = os.path.join(HERE, "dataset_sample")
        shutil.rmtree(sample, ignore_errors=True)
        os.makedirs(sample)
        for rel in sorted(manifest)[:3] + sorted(manifest)[-3:]:
            shutil.copy(os.path.join(tmp, rel), os.path.join(sample, rel.replace(os.sep, "__")))
        # end-to-end, dump only, fragment + two keyword sections
        shutil.rmtree(b.dataset_dir)
        shutil.rmtree(b.gjfdir)
        b = DatasetBuilder(atomname=list(s.atomname), dumpfilename=dump, dataset_name="golddump", cutoff=4.0,
                           stepinterval=3, n_clusters=10000, nproc=1, fragment=True, atom_pref=True,
                           qmkeywords=["%nproc=4\n#force mn15/6-31g**", "%nproc=4\n#force mn15/def2tzvp geom=chk"])
        b.builddataset()
        manifest = {}
        for base in (b.dataset_dir, b.gjfdir):
            for dirpath, _, files in os.walk(base):
                for fn in files:
                    if fn.endswith(".npy"):
                        manifest[os.path.relpath(os.path.join(dirpath, fn), tmp)] = np.load(os.path.join(dirpath, fn)).tolist()
                    else:
                        manifest[os.path.relpath(os.path.join(dirpath, fn), tmp)] = sha(os.path.join(dirpath, fn))
        g["e2e_dump"] = {"atombondtype": list(b.atombondtype), "nstructure": int(b._nstructure), "manifest": manifest}
        # end-to-end, the trajectory given as two files each (the reference strides within every file,
        # datasetbuilder.py:629-638), clustering switched off by n_clusters
        shutil.rmtree(b.dataset_dir)
        shutil.rmtree(b.gjfdir)
        parts = split_inputs(dump, bonds, dd.steplinenum, db.steplinenum, tmp)
        b = DatasetBuilder(atomname=list(s.atomname), bondfilename=parts[1], dumpfilename=parts[0], dataset_name="goldparts",
                           cutoff=4.5, stepinterval=2, n_clusters=10000, nproc=1)
        b.builddataset()
        manifest = {}
        for base in (b.dataset_dir, b.gjfdir):
            for dirpath, _, files in os.walk(base):
                for fn in files:
                    manifest[os.path.relpath(os.path.join(dirpath, fn), tmp)] = sha(os.path.join(dirpath, fn))
        g["e2e_parts"] = {"atombondtype": list(b.atombondtype), "nstructure": int(b._nstructure), "manifest": manifest}
    finally:
        os.chdir(cwd)
        shutil.rmtree(tmp, ignore_errors=True)
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(g, f, indent=0, sort_keys=True)
    print("wrote", os.path.join(HERE, "golden.json"), "structures:", g["e2e_bonds"]["nstructure"], g["e2e_dump"]["nstructure"])


if __name__ == "__main__":
    main()
