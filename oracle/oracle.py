"""CPU oracle for the per-frame analysis path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module, and only as the checker or as the
timed CPU baseline.  The product package ``mddatasetbuilder_b200`` never imports it.

What backs each function (reference file:line, all under /root/reference):

* ``connect_the_dots`` / ``crd2bond``  -> ``detect.py:249-292`` (Open Babel
  ``ConnectTheDots``; restated in ``oracle.c``; PARITY UNPINNED beyond
  ``tests/test_bond.py:9-34`` because Open Babel is not installable here).
* ``dps``                              -> ``dps.pyx:17-46`` (pinned against the
  reference's own compiled extension, ``oracle/_ref``).
* ``bond_keys_bondfile``               -> ``detect.py:105-119``.
* ``sphere`` / ``coulomb_eigs``        -> ``datasetbuilder.py:305-349`` (ASE MIC
  restated in ``oracle.c``; eigenvalues by the real ``numpy.linalg.eigh``).
* ``feature_rows``                     -> ``datasetbuilder.py:236-276`` closed form.
* ``minmax_scale`` / ``kmeans_labels`` / ``minibatch_update`` -> the real scikit-learn
  (``datasetbuilder.py:369-376``).
"""

from __future__ import annotations

import ctypes as C
import os
import subprocess
from collections import Counter

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

ATOMIC_NUMBERS = {"H": 1, "He": 2, "Li": 3, "Be": 4, "B": 5, "C": 6, "N": 7, "O": 8, "F": 9,
                  "Ne": 10, "Na": 11, "Mg": 12, "Al": 13, "Si": 14, "P": 15, "S": 16, "Cl": 17,
                  "Ar": 18}


def build(force: bool = False) -> None:
    """Compile liboracle.so (and oracle/_ref when /root/reference is present)."""
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "all"], stdout=subprocess.DEVNULL)
    elif os.path.isdir("/root/reference") and not os.path.isdir(os.path.join(_HERE, "_ref")):
        subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)


def lib():
    global _LIB
    if _LIB is None:
        build()
        L = C.CDLL(os.path.join(_HERE, "liboracle.so"))
        dp = C.POINTER(C.c_double)
        ip = C.POINTER(C.c_int)
        L.orc_rcov.restype = C.c_double
        L.orc_rcov.argtypes = [C.c_int]
        L.orc_maxbonds.restype = C.c_int
        L.orc_maxbonds.argtypes = [C.c_int]
        L.orc_free.argtypes = [C.c_void_p]
        L.orc_ob_cell_setup.argtypes = [dp, dp, dp]
        L.orc_ob_mic.argtypes = [dp, dp, dp]
        L.orc_connect_the_dots.restype = C.c_int
        L.orc_connect_the_dots.argtypes = [C.c_int, ip, dp, dp, C.c_int, C.c_int,
                                           C.POINTER(C.POINTER(C.c_int)), ip]
        L.orc_ctd_sample.restype = C.c_long
        L.orc_ctd_sample.argtypes = [C.c_int, ip, dp, dp, C.c_int, C.c_int, C.c_int]
        L.orc_bonds_to_csr.argtypes = [C.c_int, C.c_int, ip, ip, ip]
        L.orc_dps.restype = C.c_int
        L.orc_dps.argtypes = [C.c_int, ip, ip, ip, ip, ip]
        L.orc_bond_keys.restype = C.c_int
        L.orc_bond_keys.argtypes = [C.c_int, ip, ip, dp, C.POINTER(C.c_uint64)]
        L.orc_ase_dist_general.restype = C.c_double
        L.orc_ase_dist_general.argtypes = [dp, C.c_int, dp, dp]
        L.orc_sphere.restype = C.c_int
        L.orc_sphere.argtypes = [C.c_int, dp, dp, C.c_int, C.c_int, C.c_double, ip]
        L.orc_coulomb_matrix.argtypes = [C.c_int, ip, dp, dp, dp, C.c_int, dp]
        L.orc_descriptor_gather.restype = C.c_long
        L.orc_descriptor_gather.argtypes = [C.c_int, ip, dp, dp, dp, C.c_int, C.c_double, C.c_int, ip,
                                            ip, ip, C.c_long, dp, C.c_long]
        _LIB = L
    return _LIB


def _d(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _i(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


# ----------------------------------------------------------------------------
# bonds
# ----------------------------------------------------------------------------
def connect_the_dots(Z, pos, cell=None, periodic=True, mode=1):
    """Open Babel ConnectTheDots restated.  Returns (bonds[nb,2] in OBMolBondIter
    order, number of bonds deleted by the valence/angle cleanup).
    mode 0 = the O(N^2) pair loop, mode 1 = cell-list candidates (same result)."""
    Z = _i32(Z)
    pos = _f64(pos).reshape(-1, 3)
    n = len(Z)
    cell = _f64(cell if cell is not None else np.eye(3)).reshape(9)
    out = C.POINTER(C.c_int)()
    ndel = C.c_int(0)
    nb = lib().orc_connect_the_dots(n, _i(Z), _d(pos), _d(cell), int(bool(periodic)), int(mode),
                                    C.byref(out), C.byref(ndel))
    bonds = np.ctypeslib.as_array(out, shape=(max(nb, 1), 2))[:nb].copy()
    lib().orc_free(out)
    return bonds, ndel.value


def ctd_sample(Z, pos, cell, periodic, stride, offset):
    """Bounded sample of the O(N^2) ConnectTheDots pair loop (CPU-baseline timing only):
    z-ranks offset, offset+stride, ...; a full frame costs ~stride x this call."""
    Z = _i32(Z)
    pos = _f64(pos).reshape(-1, 3)
    cell = _f64(cell).reshape(9)
    return lib().orc_ctd_sample(len(Z), _i(Z), _d(pos), _d(cell), int(bool(periodic)), int(stride), int(offset))


def bonds_to_csr(n, bonds):
    """Adjacency in the append order of detect.py:282-291 as CSR (rowptr, col)."""
    bonds = _i32(bonds).reshape(-1, 2)
    rowptr = np.zeros(n + 1, dtype=np.int32)
    col = np.zeros(max(2 * len(bonds), 1), dtype=np.int32)
    lib().orc_bonds_to_csr(n, len(bonds), _i(bonds), _i(rowptr), _i(col))
    return rowptr, col[: 2 * len(bonds)]


def csr_to_lists(rowptr, col):
    return [list(map(int, col[rowptr[i]:rowptr[i + 1]])) for i in range(len(rowptr) - 1)]


def lists_to_csr(lists):
    rowptr = np.zeros(len(lists) + 1, dtype=np.int32)
    for i, l in enumerate(lists):
        rowptr[i + 1] = rowptr[i] + len(l)
    col = np.fromiter((x for l in lists for x in l), dtype=np.int32, count=int(rowptr[-1]))
    return rowptr, col


def crd2bond(Z, pos, cell, periodic, readlevel=False, mode=1):
    """``DetectDump._crd2bond`` (detect.py:249-292): adjacency lists (readlevel=False)
    or per-atom bond-order lists (readlevel=True; PerceiveBondOrders is not restated,
    every surviving bond has order 1 -- the documented tier-1 deviation)."""
    bonds, _ = connect_the_dots(Z, pos, cell, periodic, mode)
    rowptr, col = bonds_to_csr(len(Z), bonds)
    adj = csr_to_lists(rowptr, col)
    if readlevel:
        return [[1] * len(a) for a in adj]
    return adj


def canonical_csr(rowptr, col):
    """Row-sorted CSR for order-independent comparison."""
    col = np.array(col, dtype=np.int32, copy=True)
    for i in range(len(rowptr) - 1):
        col[rowptr[i]:rowptr[i + 1]].sort()
    return np.asarray(rowptr, dtype=np.int32), col


# ----------------------------------------------------------------------------
# connectivity
# ----------------------------------------------------------------------------
def dps(rowptr, col):
    """dps.pyx:17-46.  Returns (mol_atoms, mol_ptr, labels) with labels = min index."""
    rowptr = _i32(rowptr)
    col = _i32(col) if len(col) else np.zeros(1, dtype=np.int32)
    n = len(rowptr) - 1
    mol_atoms = np.zeros(max(n, 1), dtype=np.int32)
    mol_ptr = np.zeros(n + 1, dtype=np.int32)
    labels = np.zeros(max(n, 1), dtype=np.int32)
    nm = lib().orc_dps(n, _i(rowptr), _i(col), _i(mol_atoms), _i(mol_ptr), _i(labels))
    return mol_atoms[:n], mol_ptr[: nm + 1], labels[:n]


def dps_lists(bonds):
    rowptr, col = lists_to_csr(bonds)
    ma, mp, _ = dps(rowptr, col)
    return [list(map(int, ma[mp[k]:mp[k + 1]])) for k in range(len(mp) - 1)]


def ref_dps():
    """The reference's own compiled ``dps`` (oracle/_ref), or None if not built."""
    import importlib.util
    import glob

    cands = glob.glob(os.path.join(_HERE, "_ref", "mddatasetbuilder", "dps*.so"))
    if not cands:
        return None
    spec = importlib.util.spec_from_file_location("dps", cands[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.dps


# ----------------------------------------------------------------------------
# bond-type keys
# ----------------------------------------------------------------------------
def pack_key(elem, levels):
    k = ((elem + 1) << 56) | (len(levels) << 48)
    for i, l in enumerate(sorted(levels)[:12]):
        k |= min(int(l), 15) << (4 * i)
    return k


def unpack_key(key):
    key = int(key)
    elem = (key >> 56) - 1
    n = (key >> 48) & 0xFF
    return elem, [(key >> (4 * i)) & 0xF for i in range(min(n, 12))]


def key_string(key, atomname):
    """datasetbuilder.py:608-614 ``_bondtype``: e.g. 'C1111', 'H1', 'O'."""
    elem, levels = unpack_key(key)
    return f"{atomname[elem]}{''.join(map(str, levels))}"


def bond_keys_bondfile(elem, rowptr, bo):
    """detect.py:112-116: levels = sorted(max(1, round(bo)))."""
    elem = _i32(elem)
    rowptr = _i32(rowptr)
    bo = _f64(bo) if len(bo) else np.zeros(1)
    keys = np.zeros(len(elem), dtype=np.uint64)
    lib().orc_bond_keys(len(elem), _i(elem), _i(rowptr), _d(bo), keys.ctypes.data_as(C.POINTER(C.c_uint64)))
    return keys


def bond_keys_degree(elem, rowptr):
    """dump-only tier-1 keys: every bond order 1 (see crd2bond)."""
    deg = np.diff(np.asarray(rowptr, dtype=np.int64))
    return np.array([pack_key(int(e), [1] * int(d)) for e, d in zip(elem, deg)], dtype=np.uint64)


# ----------------------------------------------------------------------------
# descriptor
# ----------------------------------------------------------------------------
def coulomb_diag(symbols):
    """datasetbuilder.py:145-147."""
    return {s: ATOMIC_NUMBERS[s] ** 2.4 / 2 for s in symbols}


def sphere(pos, L, periodic, a, cutoff):
    pos = _f64(pos).reshape(-1, 3)
    L = _f64(L).reshape(3)
    out = np.zeros(len(pos), dtype=np.int32)
    m = lib().orc_sphere(len(pos), _d(pos), _d(L), int(bool(periodic)), int(a), float(cutoff), _i(out))
    return out[:m].copy()


def coulomb_matrix(Z, diag, pos, L, periodic):
    Z = _i32(Z)
    diag = _f64(diag)
    pos = _f64(pos).reshape(-1, 3)
    L = _f64(L).reshape(3)
    m = len(Z)
    M = np.zeros((m, m))
    lib().orc_coulomb_matrix(m, _i(Z), _d(diag), _d(pos), _d(L), int(bool(periodic)), _d(M))
    return M


def descriptor_gather(Z, diag_by_atom, pos, L, periodic, cutoff, targets):
    """For every target: sphere members (ascending) and Coulomb matrix.
    Returns (counts, members_list, matrices_list)."""
    Z = _i32(Z)
    pos = _f64(pos).reshape(-1, 3)
    L = _f64(L).reshape(3)
    diag_by_atom = _f64(diag_by_atom)
    targets = _i32(targets)
    nt = len(targets)
    counts = np.zeros(max(nt, 1), dtype=np.int32)
    mcap = max(nt, 1) * 64
    while True:
        members = np.zeros(mcap, dtype=np.int32)
        xcap = mcap * 64
        mats = np.zeros(xcap)
        r = lib().orc_descriptor_gather(len(Z), _i(Z), _d(diag_by_atom), _d(pos), _d(L), int(bool(periodic)),
                                        float(cutoff), nt, _i(targets), _i(counts), _i(members), mcap,
                                        _d(mats), xcap)
        if r >= 0:
            break
        mcap *= 4
    mem, mat = [], []
    mo = xo = 0
    for t in range(nt):
        m = int(counts[t])
        mem.append(members[mo:mo + m].copy())
        mat.append(mats[xo:xo + m * m].reshape(m, m).copy())
        mo += m
        xo += m * m
    return counts[:nt], mem, mat


def coulomb_eigs(mats):
    """datasetbuilder.py:349: ascending eigenvalues by numpy.linalg.eigh (LAPACK)."""
    out = [None] * len(mats)
    by_n = {}
    for i, M in enumerate(mats):
        by_n.setdefault(M.shape[0], []).append(i)
    for n, idx in by_n.items():
        w = np.linalg.eigh(np.stack([mats[i] for i in idx]))[0]
        for k, i in enumerate(idx):
            out[i] = w[k]
    return out


def feature_rows(eigs, elem_counts, pads):
    """Closed form of datasetbuilder.py:236-276: row = sort(eig U {pad_el repeated
    (max_el - cnt_el)}).  eigs: list of 1-D arrays; elem_counts [rows, n_el];
    pads [n_el] (= Z^2.4/2 per element).  Returns (X[rows, D], max_counter)."""
    elem_counts = np.asarray(elem_counts, dtype=np.int64)
    maxc = elem_counts.max(axis=0) if len(elem_counts) else np.zeros(len(pads), dtype=np.int64)
    D = int(maxc.sum())
    X = np.zeros((len(eigs), D))
    for j, e in enumerate(eigs):
        extra = np.concatenate([np.full(int(maxc[el] - elem_counts[j, el]), pads[el]) for el in range(len(pads))])
        row = np.concatenate([np.asarray(e, dtype=np.float64), extra])
        X[j] = np.sort(row)
    return X, maxc


def reference_feature_assembly(results, coulumbdiag):
    """Literal restatement of the parent-side loop datasetbuilder.py:236-275 (used to
    pin the closed form above).  results: iterable of (vector, Counter)."""
    results = list(results)
    max_counter = Counter()
    feedvector = np.zeros((len(results), 0))
    vector_elements = {}
    for j, (vector, symbols_counter) in enumerate(results):
        for element in (symbols_counter - max_counter).elements():
            vector_elements.setdefault(element, []).append(feedvector.shape[1])
            feedvector = np.pad(feedvector, ((0, 0), (0, 1)), "constant",
                                constant_values=(0, coulumbdiag[element]))
        feedvector[j, sum((vector_elements[x[0]][: x[1]] for x in symbols_counter.items()), [])] = vector
        max_counter |= symbols_counter
    return np.sort(feedvector), max_counter


# ----------------------------------------------------------------------------
# clustering: the real scikit-learn
# ----------------------------------------------------------------------------
def minmax_scale(X):
    from sklearn import preprocessing

    sc = preprocessing.MinMaxScaler()
    Xs = np.array(sc.fit_transform(X))
    return Xs, sc.data_min_, sc.data_max_, sc.scale_, sc.min_


def kmeans_labels(X, centers):
    """Assignment step of datasetbuilder.py:376 given centroids
    (sklearn.cluster._kmeans._labels_inertia -> lloyd_iter_chunked_dense)."""
    from sklearn.cluster._kmeans import _labels_inertia

    X = np.ascontiguousarray(X, dtype=np.float64)
    centers = np.ascontiguousarray(centers, dtype=np.float64)
    labels, inertia = _labels_inertia(X, np.ones(len(X)), centers, n_threads=1)
    return np.asarray(labels), float(inertia)


def minibatch_update(X, sample_weight, centers, weight_sums, labels):
    """One M-step of MiniBatchKMeans (sklearn _k_means_minibatch._minibatch_update_dense)."""
    from sklearn.cluster._k_means_minibatch import _minibatch_update_dense

    centers_new = np.empty_like(centers)
    weight_sums = np.array(weight_sums, dtype=np.float64, copy=True)
    _minibatch_update_dense(np.ascontiguousarray(X), np.ascontiguousarray(sample_weight, dtype=np.float64),
                            np.ascontiguousarray(centers), centers_new, weight_sums,
                            np.ascontiguousarray(labels, dtype=np.int32), 1)
    return centers_new, weight_sums
