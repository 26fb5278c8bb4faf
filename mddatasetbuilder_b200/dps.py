"""``dps(bonds) -> molecules``: connected molecules with the reference's exact ordering.

Drop-in for the reference's only native component, the Cython/C++ extension
``mddatasetbuilder.dps`` (dps.pyx:17-46, c_stack.cpp:1-33; signature dps.pyi:1).  The
components are found on the GPU by union-find (hook + pointer jumping) and every molecule is
then walked by one GPU thread that replays the reference's LIFO depth-first visiting order, so
the output is identical: molecules ascending in their smallest atom index, members in the order
``dps`` appends them.  The adjacency must be symmetric (LAMMPS bond tables and Open Babel bond
lists are); an asymmetric adjacency raises ``RuntimeError``.
"""

from __future__ import annotations

import numpy as np


def dps(bonds, device: int = 0):
    import torch

    from .engine import get_engine

    eng = get_engine(device, None)
    n = len(bonds)
    deg = np.fromiter((len(b) for b in bonds), dtype=np.int64, count=n)
    rowptr = np.zeros(n + 1, dtype=np.int32)
    np.cumsum(deg, out=rowptr[1:])
    col = np.fromiter((x for b in bonds for x in b), dtype=np.int32, count=int(rowptr[-1]))
    if col.size and (col.min() < 0 or col.max() >= n):
        raise RuntimeError("dps: neighbour index out of range")
    ma, mp = eng.dps(torch.as_tensor(rowptr, device=eng.device),
                     torch.as_tensor(col if col.size else np.zeros(1, dtype=np.int32), device=eng.device))
    ma = ma.cpu().numpy()
    mp = mp.cpu().numpy()
    return [ma[mp[k]:mp[k + 1]].tolist() for k in range(len(mp) - 1)]


def dps_from_ell(eng, nbr):
    """Molecules from one frame's fixed-width neighbour table [N, width] (-1 padded, row order =
    adjacency order, e.g. the file order of a LAMMPS bond table, detect.py:137-145)."""
    import torch

    nbr = np.asarray(nbr)
    n = nbr.shape[0]
    valid = nbr >= 0
    rowptr = np.zeros(n + 1, dtype=np.int32)
    np.cumsum(valid.sum(axis=1), out=rowptr[1:])
    col = nbr[valid].astype(np.int32)
    ma, mp = eng.dps(torch.as_tensor(rowptr, device=eng.device),
                     torch.as_tensor(col if col.size else np.zeros(1, dtype=np.int32), device=eng.device))
    ma = ma.cpu().numpy()
    mp = mp.cpu().numpy()
    return [ma[mp[k]:mp[k + 1]].tolist() for k in range(len(mp) - 1)]


def dps_from_ell_arrays(eng, nbr):
    """Like dps_from_ell but returns the flat (mol_atoms, mol_ptr) numpy arrays."""
    import torch

    nbr = np.asarray(nbr)
    n = nbr.shape[0]
    valid = nbr >= 0
    rowptr = np.zeros(n + 1, dtype=np.int32)
    np.cumsum(valid.sum(axis=1), out=rowptr[1:])
    col = nbr[valid].astype(np.int32)
    ma, mp = eng.dps(torch.as_tensor(rowptr, device=eng.device),
                     torch.as_tensor(col if col.size else np.zeros(1, dtype=np.int32), device=eng.device))
    return ma.cpu().numpy(), mp.cpu().numpy()
