"""Helpers shared by the parity tests."""
import numpy as np

from oracle import oracle as O


def ell_to_canonical(ell_frame):
    """ELL [N, maxb] (numpy) -> canonical (row-sorted) CSR."""
    ell_frame = np.asarray(ell_frame)
    deg = (ell_frame >= 0).sum(axis=1)
    rowptr = np.zeros(len(ell_frame) + 1, dtype=np.int32)
    np.cumsum(deg, out=rowptr[1:])
    col = ell_frame[ell_frame >= 0].astype(np.int32)  # row-major order keeps rows contiguous
    return O.canonical_csr(rowptr, col)


def oracle_frame(sys, pos, periodic=True, mode=1, cell=None):
    Z = np.array([O.ATOMIC_NUMBERS[sys.atomname[t]] for t in sys.types], dtype=np.int32)
    bonds, ndel = O.connect_the_dots(Z, pos, sys.cell() if cell is None else cell, periodic, mode)
    rowptr, col = O.bonds_to_csr(len(Z), bonds)
    return rowptr, col, ndel


def assert_ell_rows_sorted_padded(ell_frame):
    e = np.asarray(ell_frame)
    valid = e >= 0
    # padding only at the end of each row
    assert np.all(valid[:, :-1] | ~valid[:, 1:]), "holes inside ELL rows"
    a, b = e[:, :-1], e[:, 1:]
    ok = ~(valid[:, :-1] & valid[:, 1:]) | (a < b)
    assert np.all(ok), "ELL rows are not strictly ascending"
    assert np.all(e[~valid] == -1)
