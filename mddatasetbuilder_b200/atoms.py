"""A minimal ``Atoms`` container.

ASE is not a dependency of this package.  ``DetectDump._crd2bond`` (and the reference test
``tests/test_bond.py``) only touch ``len()``, ``get_atomic_numbers()``, ``.positions``,
``.pbc.any()`` and ``.cell[i][j]`` (reference detect.py:253-272), so any object exposing
those -- an ``ase.Atoms`` included -- is accepted; this class is what the package itself
returns from ``readcrd``.
"""

from __future__ import annotations

import re

import numpy as np

SYMBOLS = ["X", "H", "He", "Li", "Be", "B", "C", "N", "O", "F", "Ne", "Na", "Mg", "Al", "Si", "P", "S", "Cl",
           "Ar"]
atomic_numbers = {s: z for z, s in enumerate(SYMBOLS)}


class Atoms:
    def __init__(self, symbols=None, positions=None, cell=None, pbc=False, numbers=None, tags=None):
        if isinstance(symbols, str):
            syms = []
            for sym, cnt in re.findall(r"([A-Z][a-z]?)(\d*)", symbols):
                syms += [sym] * (int(cnt) if cnt else 1)
            numbers = [atomic_numbers[s] for s in syms]
        elif symbols is not None:
            numbers = [atomic_numbers[s] for s in symbols]
        self.numbers = np.asarray(numbers if numbers is not None else [], dtype=int)
        self.positions = (np.array(positions, dtype=float).reshape(-1, 3) if positions is not None
                          else np.zeros((len(self.numbers), 3)))
        c = np.zeros((3, 3)) if cell is None else np.array(cell, dtype=float)
        self.cell = np.diag(c) if c.shape == (3,) else c.reshape(3, 3)
        self.pbc = np.array([pbc] * 3 if isinstance(pbc, (bool, np.bool_)) else pbc, dtype=bool)
        self.tags = np.zeros(len(self.numbers), dtype=int) if tags is None else np.asarray(tags, dtype=int)

    def __len__(self):
        return len(self.numbers)

    def get_atomic_numbers(self):
        return self.numbers.copy()

    def get_chemical_symbols(self):
        return [SYMBOLS[z] for z in self.numbers]

    def get_pbc(self):
        return self.pbc.copy()

    def get_tags(self):
        return self.tags.copy()

    def __getitem__(self, i):
        i = np.asarray(list(i) if isinstance(i, range) else i)
        if i.dtype == bool:
            i = np.nonzero(i)[0]
        i = np.atleast_1d(i)
        return Atoms(numbers=self.numbers[i], positions=self.positions[i], cell=self.cell.copy(), pbc=self.pbc.copy(),
                     tags=self.tags[i])
