"""Output formatting of the selected clusters: wrap, ``.xyz``, ``.gjf`` (host side, small data).

Text formats follow the reference byte for byte: ASE's simple xyz writer as called at
datasetbuilder.py:577-585, ``Atoms.wrap`` as called at :572-576, the Gaussian input of
``_convertgjf`` (:468-521) and ``detect_multiplicity`` (:446-466).
"""

from __future__ import annotations

import itertools
import os

import numpy as np

from .atoms import atomic_numbers


def wrap_positions(positions, cell, center, pbc, eps=1e-7):
    """ASE ``wrap_positions`` (geometry.py) as ``Atoms.wrap(center=..., pbc=...)`` calls it."""
    positions = np.asarray(positions, dtype=float)
    cell = np.asarray(cell, dtype=float).reshape(3, 3)
    pbc = np.array([pbc] * 3 if isinstance(pbc, (bool, np.bool_)) else pbc, dtype=bool)
    if not hasattr(center, "__len__"):
        center = (center,) * 3
    shift = np.asarray(center, dtype=float) - 0.5 - eps
    shift[np.logical_not(pbc)] = 0.0
    fractional = np.linalg.solve(cell.T, positions.T).T - shift
    for i, periodic in enumerate(pbc):
        if periodic:
            fractional[:, i] %= 1.0
            fractional[:, i] += shift[i]
    return np.dot(fractional, cell)


class FrameWrapper:
    """``wrap_positions`` for many clusters of one frame: the cell, its transpose and the periodic axes are
    prepared once; every call then runs the same NumPy operations on arrays of the same layout as
    ``wrap_positions`` does, so the coordinates are bit-identical."""

    def __init__(self, cell, pbc, eps=1e-7):
        self.cell = np.asarray(cell, dtype=float).reshape(3, 3)
        self.cell_t = self.cell.T
        self.pbc = np.array([pbc] * 3 if isinstance(pbc, (bool, np.bool_)) else pbc, dtype=bool)
        self.free = np.logical_not(self.pbc)
        self.all_periodic = bool(self.pbc.all())
        self.eps = eps

    def __call__(self, positions, center):
        shift = np.asarray(center, dtype=float) - 0.5 - self.eps
        shift[self.free] = 0.0
        fractional = np.linalg.solve(self.cell_t, positions.T).T - shift
        if self.all_periodic:
            fractional %= 1.0
            fractional += shift
        else:
            for i in range(3):
                if self.pbc[i]:
                    fractional[:, i] %= 1.0
                    fractional[:, i] += shift[i]
        return np.dot(fractional, self.cell)


def write_xyz(path, symbols, positions):
    """ASE simple xyz: count line, empty comment line, ``%-2s %22.15f %22.15f %22.15f``."""
    with open(path, "w") as f:
        f.write("%d\n%s\n" % (len(symbols), ""))
        for s, (x, y, z) in zip(symbols, positions):
            f.write("%-2s %s %s %s\n" % (s, "%22.15f" % x, "%22.15f" % y, "%22.15f" % z))


def write_xyz_batch(paths, symbols_list, positions_list, nthreads=0):
    """The same files as ``write_xyz`` for many structures at once, formatted and written by the native
    library on several threads (csrc/textio.cpp: mdb_write_xyz_batch) -- byte-identical output."""
    import ctypes as C

    from . import _lib

    n = len(paths)
    if n == 0:
        return
    off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum([len(s) for s in symbols_list], out=off[1:])
    sym = np.concatenate([np.asarray(s, dtype="S4") for s in symbols_list])
    sym = np.ascontiguousarray(sym.astype("S4"))
    pos = np.ascontiguousarray(np.concatenate([np.asarray(p, dtype=np.float64).reshape(-1, 3) for p in positions_list]))
    arr = (C.c_char_p * n)(*[os.fsencode(p) for p in paths])
    rc = _lib.lib().mdb_write_xyz_batch(n, arr, C.c_void_p(off.ctypes.data), C.c_void_p(sym.ctypes.data),
                                        C.c_void_p(pos.ctypes.data), int(nthreads))
    if rc != 0:
        raise RuntimeError(_lib.last_error())


def write_gjf_batch(paths, molsizes_list, symbols_list, positions_list, qmkeywords, fragment, nthreads=0):
    """The same files as ``write_gjf`` for many structures at once, formatted and written by the native
    library on several threads (csrc/textio.cpp: mdb_write_gjf_batch) -- byte-identical output.
    molsizes_list[k]: sizes of the molecules of structure k (consecutive runs of its atoms)."""
    import ctypes as C

    from . import _lib

    n = len(paths)
    if n == 0:
        return
    off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum([len(s) for s in symbols_list], out=off[1:])
    moff = np.zeros(n + 1, dtype=np.int64)
    np.cumsum([len(m) for m in molsizes_list], out=moff[1:])
    msz = np.ascontiguousarray(np.concatenate([np.asarray(m, dtype=np.int32) for m in molsizes_list]), dtype=np.int32)
    flat = np.concatenate([np.asarray(s) for s in symbols_list])
    sym = np.ascontiguousarray(flat.astype("S4"))
    names, inverse = np.unique(flat, return_inverse=True)
    znum = np.ascontiguousarray(np.array([atomic_numbers[str(x)] for x in names], dtype=np.int32)[inverse])
    pos = np.ascontiguousarray(np.concatenate([np.asarray(p, dtype=np.float64).reshape(-1, 3) for p in positions_list]))
    arr = (C.c_char_p * n)(*[os.fsencode(p) for p in paths])
    kws = [k.encode() for k in qmkeywords]
    karr = (C.c_char_p * len(kws))(*kws)
    rc = _lib.lib().mdb_write_gjf_batch(n, arr, C.c_void_p(off.ctypes.data), C.c_void_p(sym.ctypes.data),
                                        C.c_void_p(znum.ctypes.data), C.c_void_p(pos.ctypes.data),
                                        C.c_void_p(moff.ctypes.data), C.c_void_p(msz.ctypes.data), len(kws), karr,
                                        int(bool(fragment)), int(nthreads))
    if rc != 0:
        raise RuntimeError(_lib.last_error())


def detect_multiplicity(symbols):
    """Reference datasetbuilder.py:446-466 (charge 0; O2 is a triplet)."""
    if list(symbols) == ["O", "O"]:
        return 3
    n_total = sum(atomic_numbers[s] for s in symbols)
    return n_total % 2 + 1


def write_gjf(gjffilename, takenatomidindex, symbols, positions, qmkeywords, fragment):
    """Reference ``_convertgjf`` (datasetbuilder.py:468-521).  takenatomidindex: list of
    ranges, one per molecule, into the cluster's atom order."""
    symbols = list(symbols)
    buff = []
    multiplicities = [detect_multiplicity([symbols[i] for i in atoms]) for atoms in takenatomidindex]
    multiplicity_whole = sum(multiplicities) - len(takenatomidindex) + 1
    multiplicity_whole_str = f"0 {multiplicity_whole}"
    title = "\nGenerated by MDDatasetMaker (Author: Jinzhe Zeng)\n"
    connect = None
    if len(qmkeywords) > 1:
        connect = "\n--link1--\n"
        chk = [f"%chk={os.path.splitext(os.path.basename(gjffilename))[0]}.chk"]
    else:
        chk = []
    if len(takenatomidindex) == 1 or not fragment:
        buff.extend((*chk, qmkeywords[0], title, multiplicity_whole_str))
        buff.extend("{} {:.5f} {:.5f} {:.5f}".format(s, *p) for s, p in zip(symbols, positions))
    else:
        kw0 = f"{qmkeywords[0]} guess=fragment={len(takenatomidindex)}"
        multiplicities_str = "{} {}".format(multiplicity_whole_str,
                                            " ".join([f"0 {multiplicity}" for multiplicity in multiplicities]))
        buff.extend((*chk, kw0, title, multiplicities_str))
        for index, atoms in enumerate(takenatomidindex, 1):
            buff.extend(["{}(Fragment={}) {:.5f} {:.5f} {:.5f}".format(symbols[i], index, *positions[i]) for i in atoms])
    for kw in itertools.islice(qmkeywords, 1, None):
        buff.extend((connect, *chk, kw, title, f"0 {multiplicity_whole}", "\n"))
    buff.append("\n")
    with open(gjffilename, "w") as f:
        f.write("\n".join(buff))
