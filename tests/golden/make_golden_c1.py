"""C1-sized golden run (~300 atoms, 200 frames, `-i 10 -s 10`, mirroring the reference's
tests/test.json:17-23) through the reference's OWN, unmodified Python, capturing what its parent loop
builds per bond type: the sorted `feedvector` and `stepatom` (reference datasetbuilder.py:236-276), the
bond types in first-seen order, and the max counters.

Run in the build container only (needs /root/reference; `make -C oracle` first):

    python tests/golden/make_golden_c1.py

`DatasetBuilder._clusterdatas` is replaced for the run by a recorder (the clustering itself is
scikit-learn's, unseeded in the reference, and is pinned separately against the real scikit-learn); every
other line executed is the reference's.  The trajectory text is not committed: tests/util.py::c1_inputs
regenerates it from the seeded generator and the fixture pins its sha256.

Output: tests/golden/c1_feedvectors.npz (float64 rows per type, int64 stepatom) + c1_golden.json.
"""
import hashlib
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import refrun  # noqa: E402
from tests.util import c1_inputs  # noqa: E402


def sha(path):
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def main():
    refrun.load()
    from mddatasetbuilder.datasetbuilder import DatasetBuilder

    tmp = tempfile.mkdtemp()
    cwd = os.getcwd()
    os.chdir(tmp)
    try:
        s, dump, bonds, err = c1_inputs(tmp)
        captured = []

        def recorder(cls, X, n_clusters, n_each=1):
            captured.append(np.array(X, dtype=np.float64, copy=True))
            return np.arange(min(n_clusters, len(X)))

        orig = DatasetBuilder._clusterdatas
        DatasetBuilder._clusterdatas = classmethod(recorder)
        out = {"natoms": int(s.natoms), "nframes": 200, "stepinterval": 10, "n_clusters": 10,
               "dump_sha256": sha(dump), "bonds_sha256": sha(bonds), "runs": {}}
        arrays = {}
        try:
            for mode, bf in (("bonds", bonds), ("dump", None)):
                captured.clear()
                b = DatasetBuilder(atomname=list(s.atomname), bondfilename=bf, dumpfilename=dump, dataset_name=f"c1{mode}",
                                   cutoff=5.0, stepinterval=10, n_clusters=10, nproc=1)
                # the reference logs "Max counter of <type>"; capture the rows in the order the types are clustered
                b.trajatom_dir = tempfile.mkdtemp(dir=tmp)
                b._readtimestepsbond()
                types = list(b.atombondtype)
                clustered = []
                with open(os.path.join(b.trajatom_dir, "chooseatoms"), "wb") as fc:
                    for t in types:
                        n0 = len(captured)
                        b._writecoulumbmatrix(t, fc)
                        if len(captured) > n0:
                            clustered.append(t)
                assert len(clustered) == len(captured)
                for t, X in zip(clustered, captured):
                    arrays[f"{mode}__{t}"] = X
                out["runs"][mode] = {"atombondtype": types, "clustered": clustered, "nstep": int(b._nstep),
                                     "rows": {t: int(len(X)) for t, X in zip(clustered, captured)},
                                     "D": {t: int(X.shape[1]) for t, X in zip(clustered, captured)}}
        finally:
            DatasetBuilder._clusterdatas = orig
        np.savez_compressed(os.path.join(HERE, "c1_feedvectors.npz"), **arrays)
        with open(os.path.join(HERE, "c1_golden.json"), "w") as f:
            json.dump(out, f, indent=0, sort_keys=True)
        print("wrote c1 goldens:", {m: (len(r["atombondtype"]), len(r["clustered"])) for m, r in out["runs"].items()},
              os.path.getsize(os.path.join(HERE, "c1_feedvectors.npz")), "bytes")
    finally:
        os.chdir(cwd)


if __name__ == "__main__":
    main()
