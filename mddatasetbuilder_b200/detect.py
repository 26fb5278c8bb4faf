"""Detect bond types and molecules from LAMMPS trajectories on the GPU.

Same classes, constructor arguments, method names, argument meaning and error behaviour as the
reference's ``mddatasetbuilder/detect.py``; the per-frame arithmetic runs in the sm_100a kernels
behind ``libmdb_b200.so``:

===========================  =====================================  ===========================
method                       reference                              kernel path
===========================  =====================================  ===========================
``DetectDump._crd2bond``     detect.py:249-292 (Open Babel)         cell list + k_bonds + k_cleanup
``*.readatombondtype``       detect.py:90-119, 196-226              k_keys / k_keys_bo
``*.readmolecule``           detect.py:121-145, 228-247 (``dps``)   union-find + DFS replay
``DetectDump.readcrd``       detect.py:294-345                      C++ parser (textio.cpp)
===========================  =====================================  ===========================

Open Babel's ``PerceiveBondOrders`` (dump-only keys, ``_crd2bond(atoms, readlevel=True)``) is the
C/H/O-subset restatement of csrc/perceive.cu / oracle/perceive.py: ring passes and most
bondtyp.txt patterns are left out, parity with Open Babel itself is unpinned; the
reference's only known answer for it (tests/test_bond.py:19-20) holds.
"""

from __future__ import annotations

import os
import pickle
from abc import ABCMeta, abstractmethod
from collections import defaultdict
from enum import Enum, auto
from typing import List, Optional, Tuple

import numpy as np

from .atoms import SYMBOLS, Atoms, atomic_numbers
from .reader import KIND_BOND, KIND_DUMP, TextReader


def _device() -> int:
    return int(os.environ.get("LOCAL_RANK", "0"))


def _engine(atomname=None):
    from .engine import get_engine

    return get_engine(_device(), atomname)


def unpack_key(key: int):
    """(element index, sorted levels) from the 64-bit key written by k_keys / k_keys_bo."""
    key = int(key)
    n = (key >> 48) & 0xFF
    return (key >> 56) - 1, [(key >> (4 * i)) & 0xF for i in range(min(n, 12))]


class Detect(metaclass=ABCMeta):
    """Detect structures from file(s)."""

    def __init__(self, filename, atomname, pbc, errorlimit=None, errorfilename=None):
        self.filename = filename
        self.atomname = atomname
        self.pbc = pbc
        self.errorlimit = errorlimit
        self.errorfilename = errorfilename
        self.steplinenum = self._readN()

    @abstractmethod
    def _readN(self) -> int:
        pass

    @abstractmethod
    def readatombondtype(self, item) -> Tuple[dict, int]:
        """Read bond types of atoms such as C1111."""

    @abstractmethod
    def readmolecule(self, lines) -> Tuple[List[List[int]], Optional[Atoms]]:
        """Read molecules."""

    @staticmethod
    def gettype(inputtype):
        """Get the class for the input file type."""
        if inputtype == "bond":
            detectclass = DetectBond
        elif inputtype == "dump":
            detectclass = DetectDump
        else:
            raise RuntimeError("Wrong input file type")
        return detectclass

    # shared: {pickled (element, levels) -> [atom ids, 1-based]} from per-atom 64-bit keys
    def _group_keys(self, keys: np.ndarray, keep: Optional[np.ndarray] = None) -> dict:
        d = defaultdict(list)
        names = self.atomnames
        uniq, inverse = np.unique(keys, return_inverse=True)
        first = np.full(len(uniq), len(keys), dtype=np.int64)
        np.minimum.at(first, inverse, np.arange(len(keys)))
        for u in np.argsort(first, kind="stable"):  # dict insertion order = first atom carrying the key
            idx = np.nonzero(inverse == u)[0]
            if keep is not None:
                idx = idx[keep[idx]]
            if idx.size == 0:
                continue
            _, levels = unpack_key(uniq[u])
            d[pickle.dumps((names[idx[0]], levels))] = (idx + 1).tolist()
        return d


class DetectBond(Detect):
    """Detect from the LAMMPS bond file."""

    def _readN(self):
        """Read bondfile N, which should be at very beginning."""
        self.reader = TextReader(KIND_BOND, self.filename)
        self._N = self.reader.natoms
        self.atomtype = self.reader.atomtype.astype(int)
        self.atomnames = np.asarray(self.atomname)[self.atomtype - 1]
        return self.reader.steplinenum

    def _parse(self, lines):
        eng = _engine(self.atomname)
        return eng, self.reader.parse_lines(lines, width=eng.max_bonds)

    def readatombondtype(self, item):
        """Read bond orders of atoms.  item = ((step, lines), _); returns (dict, step)."""
        (step, lines), _ = item
        eng, (nbr, bo) = self._parse(lines)
        import torch

        keys = eng.bond_keys_bo(torch.as_tensor(nbr[None], device=eng.device), torch.as_tensor(bo[None], device=eng.device),
                                eng.to_device_types(self.atomtype - 1))
        return self._group_keys(keys.cpu().numpy().view(np.uint64)[0]), step

    def readmolecule(self, lines):
        """Return (molecules, None) from the lines of one bond-file frame."""
        eng, (nbr, _) = self._parse(lines)
        from .dps import dps_from_ell

        return dps_from_ell(eng, nbr), None


class DetectDump(Detect):
    """Detect from the dump file."""

    def _readN(self):
        self.reader = TextReader(KIND_DUMP, self.filename)
        self._N = self.reader.natoms
        self.atomtype = self.reader.atomtype.astype(int)
        self.atomnames = np.asarray(self.atomname)[self.atomtype - 1]
        self.id_idx, self.tidx, self.xidx, self.yidx, self.zidx = self.reader.cols
        return self.reader.steplinenum

    def readatombondtype(self, item):
        """Read bond orders of atoms.  item = ((step, lines), needlerror)."""
        (step, lines), needlerror = item
        keep = None
        if needlerror:
            # evident intent of detect.py:215-223 (the reference raises here at this snapshot):
            # keep the atoms whose model deviation exceeds errorlimit
            trajline, errorline = lines
            lines = trajline
            lerror = np.array(errorline.split(), dtype=float)[7:]
            keep = lerror > (self.errorlimit if self.errorlimit is not None else -np.inf)
        step_atoms, _ = self.readcrd(lines)
        level = self._crd2bond(step_atoms, readlevel=True)
        d = defaultdict(list)
        for i, (n, l) in enumerate(zip(self.atomnames, level)):
            if keep is None or (i < len(keep) and keep[i]):
                d[pickle.dumps((n, sorted(l)))].append(i + 1)
        return d, step

    def readmolecule(self, lines):
        """Return (molecules, step_atoms) from the lines of one dump frame."""
        from .dps import dps as connectmolecule

        step_atoms, _ = self.readcrd(lines)
        bond = self._crd2bond(step_atoms, readlevel=False)
        molecules = connectmolecule(bond)
        return molecules, step_atoms

    @classmethod
    def _crd2bond(cls, step_atoms, readlevel):
        """Adjacency lists (readlevel=False) or per-atom bond-order lists (readlevel=True)
        in Open Babel's bond-iteration order, exactly the shape detect.py:278-292 builds."""
        atomnumber = len(step_atoms)
        if atomnumber == 0:
            return []
        numbers = np.asarray(step_atoms.get_atomic_numbers(), dtype=int)
        zs = sorted(set(int(z) for z in numbers))
        try:
            names = [SYMBOLS[z] for z in zs]
        except IndexError:
            raise RuntimeError("element outside the built-in table (Z = 1..18)") from None
        eng = _engine(names)
        lut = {z: k for k, z in enumerate(zs)}
        types = np.array([lut[int(z)] for z in numbers], dtype=np.uint8)
        periodic = bool(np.asarray(step_atoms.pbc).any())
        cell = np.array([[step_atoms.cell[i][j] for j in range(3)] for i in range(3)], dtype=float)
        out = eng.analyze(np.asarray(step_atoms.positions, dtype=float)[None], cell.reshape(9), types, periodic,
                          want_keys=False, want_labels=False)
        eng.check()
        rowptr, col = eng.ell_to_csr(out["ell"], order=1, pos=out["pos"])
        rowptr = rowptr.cpu().numpy()[0]
        col = col.cpu().numpy()[0]
        if readlevel:
            orders = eng.perceive(out["pos"], cell.reshape(9), types, out["ell"], periodic, want_keys=False)["orders"]
            eng.check()
            orders = orders.cpu().numpy()[0]
            ell = out["ell"].cpu().numpy()[0]
            level = []
            for i in range(atomnumber):
                by_partner = {int(j): int(o) for j, o in zip(ell[i], orders[i]) if j >= 0}
                level.append([by_partner[int(j)] for j in col[rowptr[i]:rowptr[i + 1]]])
            return level
        return [col[rowptr[i]:rowptr[i + 1]].tolist() for i in range(atomnumber)]

    def readcrd(self, item):
        """Only this function can read coordinates: (Atoms sorted by id, ids)."""
        pos, cell = self.reader.parse_lines(item)
        numbers = np.array([atomic_numbers[s] for s in self.atomnames], dtype=int)
        step_atoms = Atoms(numbers=numbers, positions=pos, cell=cell, pbc=self.pbc)
        return step_atoms, list(range(1, self._N + 1))

    class LineType(Enum):
        """Line type in the LAMMPS dump files."""

        TIMESTEP = auto()
        ATOMS = auto()
        NUMBER = auto()
        BOX = auto()
        OTHER = auto()

        @classmethod
        def linecontent(cls, line):
            """Return line content."""
            if line.startswith("ITEM: TIMESTEP"):
                return cls.TIMESTEP
            if line.startswith("ITEM: ATOMS"):
                return cls.ATOMS
            if line.startswith("ITEM: NUMBER OF ATOMS"):
                return cls.NUMBER
            if line.startswith("ITEM: BOX"):
                return cls.BOX
            return cls.OTHER
