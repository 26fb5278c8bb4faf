"""ase.Atom / ase.Atoms stand-in (numpy restatement of ASE geometry).

Test infrastructure only -- PARITY UNPINNED (ASE itself is not installable here)."""
import itertools
import re

import numpy as np

from .data import atomic_numbers, chemical_symbols


class Atom:
    def __init__(self, symbol="X", position=(0, 0, 0), tag=0, atoms=None, index=None):
        self._atoms = atoms
        self._index = index
        if atoms is None:
            self._symbol = symbol
            self._position = np.array(position, dtype=float)
            self._tag = tag

    @property
    def symbol(self):
        return self._symbol if self._atoms is None else chemical_symbols[self._atoms.numbers[self._index]]

    @property
    def number(self):
        return atomic_numbers[self.symbol]

    @property
    def position(self):
        return self._position if self._atoms is None else self._atoms.positions[self._index]

    @property
    def tag(self):
        return self._tag if self._atoms is None else int(self._atoms._tags[self._index])

    @tag.setter
    def tag(self, v):
        if self._atoms is None:
            self._tag = v
        else:
            self._atoms._tags[self._index] = v


def _parse_formula(s):
    out = []
    for sym, cnt in re.findall(r"([A-Z][a-z]?)(\d*)", s):
        out += [sym] * (int(cnt) if cnt else 1)
    return out


def _complete_cell(cell):
    return np.array(cell, dtype=float).reshape(3, 3)


def naive_find_mic(v, cell):
    f = np.linalg.solve(cell.T, v.T).T
    f -= np.floor(f + 0.5)
    vmin = f @ cell
    vlen = np.linalg.norm(vmin, axis=1)
    return vmin, vlen


def general_find_mic(v, cell, pbc):
    # ASE Minkowski-reduces the cell first; for the orthorhombic / mildly tilted
    # LAMMPS cells used here the reduced cell spans the same images when the image
    # range is widened to [-2, 2] for tilted cells.
    rcell = cell
    fractional = np.linalg.solve(rcell.T, np.asarray(v).T).T
    for i in range(3):
        if pbc[i]:
            fractional[:, i] %= 1.0
    positions = fractional @ rcell
    tilted = bool(np.abs(rcell - np.diag(np.diag(rcell))).max() > 0)
    r = 2 if tilted else 1
    ranges = [np.arange(-r * p, r * p + 1) for p in pbc]
    hkls = np.array([(0, 0, 0)] + list(itertools.product(*ranges)))
    vrvecs = hkls @ rcell
    x = positions + vrvecs[:, None]
    lengths = np.linalg.norm(x, axis=2)
    indices = np.argmin(lengths, axis=0)
    vmin = x[indices, np.arange(len(positions)), :]
    vlen = lengths[indices, np.arange(len(positions))]
    return vmin, vlen


def find_mic(v, cell, pbc):
    cell = _complete_cell(cell)
    pbc = cell.any(1) & np.asarray(pbc, dtype=bool)
    dim = int(np.sum(pbc))
    v = np.atleast_2d(np.asarray(v, dtype=float))
    if dim > 0:
        safe = False
        if dim == 3:
            vmin, vlen = naive_find_mic(v, cell)
            if (vlen < 0.5 * min(np.linalg.norm(cell, axis=1))).all():
                safe = True
        if not safe:
            vmin, vlen = general_find_mic(v, cell, pbc)
    else:
        vmin = v.copy()
        vlen = np.linalg.norm(vmin, axis=1)
    return vmin, vlen


class Atoms:
    def __init__(self, symbols=None, positions=None, cell=None, pbc=None, numbers=None, tags=None):
        if isinstance(symbols, str):
            syms = _parse_formula(symbols)
            self.numbers = np.array([atomic_numbers[s] for s in syms], dtype=int)
            self.positions = np.array(positions, dtype=float).reshape(-1, 3)
        elif symbols is not None and len(symbols) and isinstance(symbols[0], Atom):
            self.numbers = np.array([a.number for a in symbols], dtype=int)
            self.positions = np.array([a.position for a in symbols], dtype=float).reshape(-1, 3)
            tags = [a.tag for a in symbols]
        elif numbers is not None:
            self.numbers = np.array(numbers, dtype=int)
            self.positions = np.array(positions, dtype=float).reshape(-1, 3)
        else:
            syms = list(symbols) if symbols is not None else []
            self.numbers = np.array([atomic_numbers[s] for s in syms], dtype=int)
            self.positions = (np.array(positions, dtype=float).reshape(-1, 3) if positions is not None
                              else np.zeros((len(syms), 3)))
        self.cell = _complete_cell(cell if cell is not None else np.zeros((3, 3)))
        if pbc is None:
            pbc = False
        self.pbc = np.array([pbc] * 3 if isinstance(pbc, (bool, np.bool_)) else pbc, dtype=bool)
        self._tags = np.array(tags if tags is not None else np.zeros(len(self.numbers)), dtype=int)

    def __len__(self):
        return len(self.numbers)

    def __iter__(self):
        for i in range(len(self)):
            yield Atom(atoms=self, index=i)

    def __getitem__(self, i):
        if isinstance(i, (int, np.integer)):
            n = len(self)
            if i < -n or i >= n:
                raise IndexError("Index out of range.")
            return Atom(atoms=self, index=int(i) % n)
        if isinstance(i, range):
            i = np.array(list(i), dtype=int)
        i = np.asarray(i)
        if i.dtype == bool:
            i = np.nonzero(i)[0]
        return Atoms(numbers=self.numbers[i], positions=self.positions[i], cell=self.cell.copy(),
                     pbc=self.pbc.copy(), tags=self._tags[i])

    def get_atomic_numbers(self):
        return self.numbers.copy()

    def get_chemical_symbols(self):
        return [chemical_symbols[z] for z in self.numbers]

    def get_tags(self):
        return self._tags.copy()

    def get_pbc(self):
        return self.pbc.copy()

    def get_cell_lengths_and_angles(self):
        lengths = np.linalg.norm(self.cell, axis=1)
        angles = []
        for i, j in ((1, 2), (0, 2), (0, 1)):
            ll = lengths[i] * lengths[j]
            angles.append(np.degrees(np.arccos(np.dot(self.cell[i], self.cell[j]) / ll)) if ll > 1e-16 else 90.0)
        return np.array(list(lengths) + angles)

    def get_distances(self, a, indices, mic=False, vector=False):
        R = self.positions
        p1 = R[[a]]
        p2 = R[list(indices) if not isinstance(indices, np.ndarray) else indices]
        D = (p2[np.newaxis, :, :] - p1[:, np.newaxis, :]).reshape((-1, 3))
        if mic:
            D, D_len = find_mic(D, self.cell, self.pbc)
        else:
            D_len = np.linalg.norm(D, axis=1)
        return D if vector else D_len

    def get_all_distances(self, mic=False, vector=False):
        R = self.positions
        n = len(R)
        ind1, ind2 = np.triu_indices(n, k=1)
        D = R[ind2] - R[ind1]
        if mic and len(D):
            D, D_len = find_mic(D, self.cell, self.pbc)
        else:
            D_len = np.linalg.norm(D, axis=1) if len(D) else np.zeros(0)
        out = np.zeros((n, n))
        out[(ind1, ind2)] = D_len
        out += out.T
        return out

    def wrap(self, center=(0.5, 0.5, 0.5), pbc=None, eps=1e-7):
        if pbc is None:
            pbc = self.pbc
        pbc = np.array([pbc] * 3 if isinstance(pbc, (bool, np.bool_)) else pbc, dtype=bool)
        if not hasattr(center, "__len__"):
            center = (center,) * 3
        shift = np.asarray(center, dtype=float) - 0.5 - eps
        shift[np.logical_not(pbc)] = 0.0
        fractional = np.linalg.solve(self.cell.T, np.asarray(self.positions).T).T - shift
        for i, periodic in enumerate(pbc):
            if periodic:
                fractional[:, i] %= 1.0
                fractional[:, i] += shift[i]
        self.positions[:] = np.dot(fractional, self.cell)
