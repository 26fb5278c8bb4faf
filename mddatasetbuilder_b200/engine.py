"""One GPU's analysis engine: thin Python wrappers over the C ABI (include/mdb_b200.h).

PyTorch supplies device memory and streams only; every computation on the path is a
hand-written sm_100a kernel inside libmdb_b200.so.
"""

from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

ATOMIC_NUMBERS = {"H": 1, "He": 2, "Li": 3, "Be": 4, "B": 5, "C": 6, "N": 7, "O": 8, "F": 9, "Ne": 10,
                  "Na": 11, "Mg": 12, "Al": 13, "Si": 14, "P": 15, "S": 16, "Cl": 17, "Ar": 18}


def _torch():
    import torch

    return torch


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _np_ptr(a):
    return C.c_void_p(a.ctypes.data) if a is not None else C.c_void_p(0)


class Engine:
    """Owns one ``mdb_ctx`` (one per process and GPU, like one Pool worker of the
    reference, utils.py:198-231, but driving a whole device)."""

    MAXT = 16   # element types per trajectory (MDB_MAXT in csrc/common.cuh)

    def __init__(self, device: int = 0, atomname=("C", "H", "O"), max_bonds: int = 8):
        torch = _torch()
        if not torch.cuda.is_available():
            raise RuntimeError("mddatasetbuilder_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.L = _lib.lib()
        self.device = torch.device("cuda", device)
        torch.cuda.set_device(self.device)
        ctx = C.c_void_p()
        _lib.check(self.L.mdb_create(int(device), C.byref(ctx)), "mdb_create")
        self.ctx = ctx
        self.max_bonds = 8
        self.set_elements(atomname)
        if max_bonds != 8:
            self.set_option("max_bonds", max_bonds)

    def close(self):
        if getattr(self, "ctx", None):
            self.L.mdb_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ config
    def set_elements(self, atomname):
        self.atomname = [str(s) for s in atomname]
        try:
            z = [ATOMIC_NUMBERS[s] for s in self.atomname]
        except KeyError as e:
            raise RuntimeError(f"element {e} is outside the built-in table (Z = 1..18)") from None
        arr = (C.c_int * len(z))(*z)
        _lib.check(self.L.mdb_set_elements(self.ctx, len(z), arr), "mdb_set_elements")
        self.Z = np.array(z, dtype=np.int32)

    def set_option(self, key: str, value: float):
        _lib.check(self.L.mdb_set_option(self.ctx, key.encode(), float(value)), f"mdb_set_option({key})")
        if key == "max_bonds":
            self.max_bonds = int(value)

    @property
    def launches(self) -> int:
        return int(self.L.mdb_launch_count(self.ctx))

    def profile(self, on: bool):
        _lib.check(self.L.mdb_profile_enable(self.ctx, int(on)), "mdb_profile_enable")
        self._profiling = bool(on)

    def profile_read(self):
        """{kernel name: (launch count, total ms)} since the last read (synchronises)."""
        buf = C.create_string_buffer(1 << 16)
        _lib.check(self.L.mdb_profile_read(self.ctx, buf, len(buf)), "mdb_profile_read")
        out = {}
        for line in buf.value.decode().splitlines():
            name, cnt, ms = line.rsplit(None, 2)
            out[name] = (int(cnt), float(ms))
        return out

    def fp64_peak_tflops(self) -> float:
        """Measured DFMA peak of this GPU (TFLOP/s)."""
        v = C.c_double(0.0)
        _lib.check(self.L.mdb_fp64_peak(self.ctx, C.byref(v), self.stream()), "mdb_fp64_peak")
        return float(v.value)

    def stream(self):
        return C.c_void_p(_torch().cuda.current_stream(self.device).cuda_stream)

    # ------------------------------------------------------------------ helpers
    def _cells(self, cell, nframes):
        if cell is None:
            return None
        c = np.ascontiguousarray(np.asarray(cell, dtype=np.float64))
        if c.size == 3:
            c = np.diag(c.reshape(3))
        if c.size == 9:
            c = np.broadcast_to(c.reshape(1, 9), (nframes, 9))
        c = np.ascontiguousarray(c.reshape(nframes, 9))
        return c

    def to_device_pos(self, pos):
        torch = _torch()
        if isinstance(pos, np.ndarray):
            pos = torch.from_numpy(np.ascontiguousarray(pos, dtype=np.float64))
        pos = pos.to(self.device, dtype=torch.float64).contiguous()
        if pos.dim() == 2:
            pos = pos[None]
        return pos

    def to_device_types(self, types):
        torch = _torch()
        if isinstance(types, np.ndarray):
            types = torch.from_numpy(np.ascontiguousarray(types).astype(np.uint8))
        return types.to(self.device, dtype=torch.uint8).contiguous()

    # ------------------------------------------------------------------ path
    def analyze(self, pos, cell, types, periodic=True, want_ell=True, want_keys=True, want_labels=True,
                out=None):
        """cell list -> bonds -> cleanup -> keys -> molecule labels for a batch of frames.

        pos [F,N,3] float64 (device tensor or numpy), cell [F,9]/[9]/[3] host, types [N]
        0-based element index.  Returns dict of device tensors: ell [F,N,max_bonds] int32,
        keys [F,N] int64 (bit pattern of the uint64 key), labels [F,N] int32."""
        torch = _torch()
        pos = self.to_device_pos(pos)
        types = self.to_device_types(types)
        F, N = int(pos.shape[0]), int(pos.shape[1])
        cells = self._cells(cell, F) if periodic else None
        out = out or {}
        ell = out.get("ell") if want_ell else None
        keys = out.get("keys") if want_keys else None
        labels = out.get("labels") if want_labels else None
        if want_ell and ell is None:
            ell = torch.empty((F, N, self.max_bonds), dtype=torch.int32, device=self.device)
        if want_keys and keys is None:
            keys = torch.empty((F, N), dtype=torch.int64, device=self.device)
        if want_labels and labels is None:
            labels = torch.empty((F, N), dtype=torch.int32, device=self.device)
        rc = self.L.mdb_analyze_frames(self.ctx, F, N, _ptr(pos), _np_ptr(cells), _ptr(types), int(bool(periodic)),
                                       _ptr(ell), _ptr(keys), _ptr(labels), self.stream())
        _lib.check(rc, "mdb_analyze_frames")
        return {"ell": ell, "keys": keys, "labels": labels, "pos": pos, "types": types}

    def check(self):
        """Synchronise and raise if a kernel flagged an error (cell/bond overflow ...)."""
        v = C.c_longlong()
        e = C.c_longlong()
        o = C.c_longlong()
        _lib.check(self.L.mdb_bond_stats(self.ctx, C.byref(v), C.byref(e), C.byref(o)), "device check")
        return {"violators": v.value, "exact_tests": e.value, "overflow": o.value}

    def perceive(self, pos, cell, types, ell, periodic=True, labels=None, want_keys=True):
        """Bond orders of the final bonds (Open Babel PerceiveBondOrders, reference detect.py:279-280,
        C/H/O subset): orders [F,N,max_bonds] uint8 per ELL slot (0 = empty) and, optionally, the
        64-bit keys built from them."""
        torch = _torch()
        pos = self.to_device_pos(pos)
        types = self.to_device_types(types)
        F, N = int(pos.shape[0]), int(pos.shape[1])
        cells = self._cells(cell, F) if periodic else None
        orders = torch.empty((F, N, self.max_bonds), dtype=torch.uint8, device=self.device)
        keys = torch.empty((F, N), dtype=torch.int64, device=self.device) if want_keys else None
        rc = self.L.mdb_perceive_bond_orders(self.ctx, F, N, _ptr(pos), _np_ptr(cells), _ptr(types), int(bool(periodic)),
                                             _ptr(ell), _ptr(labels), _ptr(orders), _ptr(keys), self.stream())
        _lib.check(rc, "mdb_perceive_bond_orders")
        return {"orders": orders, "keys": keys}

    def perceive_stats(self):
        v = C.c_longlong()
        _lib.check(self.L.mdb_perceive_stats(self.ctx, C.byref(v)), "mdb_perceive_stats")
        return {"rings": v.value}

    def ell_to_csr(self, ell, order=0, pos=None, colcap=None):
        torch = _torch()
        F, N, _ = ell.shape
        colcap = int(colcap or max(1, N * self.max_bonds))
        rowptr = torch.empty((F, N + 1), dtype=torch.int32, device=self.device)
        col = torch.full((F, colcap), -1, dtype=torch.int32, device=self.device)
        rc = self.L.mdb_ell_to_csr(self.ctx, int(F), int(N), _ptr(ell), int(order), _ptr(pos), _ptr(rowptr), _ptr(col),
                                   colcap, self.stream())
        _lib.check(rc, "mdb_ell_to_csr")
        return rowptr, col

    def bond_keys_bo(self, ell, bo, types):
        torch = _torch()
        F, N, _ = ell.shape
        keys = torch.empty((F, N), dtype=torch.int64, device=self.device)
        rc = self.L.mdb_bond_keys_bo(self.ctx, int(F), int(N), _ptr(ell), _ptr(bo), _ptr(types), _ptr(keys),
                                     self.stream())
        _lib.check(rc, "mdb_bond_keys_bo")
        return keys

    def bond_keys_csr(self, rowptr, bo, types, colcap=None):
        """Bond-file mode keys from a CSR bond table: rowptr [F,N+1] int32 (frame-local), bo [F,colcap] float64."""
        torch = _torch()
        F, N = int(rowptr.shape[0]), int(rowptr.shape[1]) - 1
        colcap = int(colcap if colcap is not None else bo.shape[1])
        keys = torch.empty((F, N), dtype=torch.int64, device=self.device)
        rc = self.L.mdb_bond_keys_csr(self.ctx, F, N, _ptr(rowptr), _ptr(bo), colcap, _ptr(types), _ptr(keys), self.stream())
        _lib.check(rc, "mdb_bond_keys_csr")
        return keys

    def components_csr(self, rowptr, col, colcap=None, out=None):
        """Molecule labels (smallest atom index of the molecule) from a CSR bond table."""
        torch = _torch()
        F, N = int(rowptr.shape[0]), int(rowptr.shape[1]) - 1
        colcap = int(colcap if colcap is not None else col.shape[1])
        labels = out if out is not None else torch.empty((F, N), dtype=torch.int32, device=self.device)
        rc = self.L.mdb_components_csr(self.ctx, F, N, _ptr(rowptr), _ptr(col), colcap, _ptr(labels), self.stream())
        _lib.check(rc, "mdb_components_csr")
        return labels

    # ------------------------------------------------------------------ group-by (streaming pass A)
    def key_table(self, cap=1024):
        """Empty device key dictionary for `group_keys`: (keys, counts, first) int64 [cap]."""
        torch = _torch()
        return (torch.full((cap,), -1, dtype=torch.int64, device=self.device),
                torch.zeros(cap, dtype=torch.int64, device=self.device),
                torch.full((cap,), -1, dtype=torch.int64, device=self.device))

    def group_keys(self, keys, table, step0=0, keep=None, want_order=True):
        """Adds a batch of keys [F,N] to the dictionary `table`; returns (codes [F,N] int16 bit pattern of the
        uint16 slot, order [F*N] int32 or None, seg [cap+1] int64 or None) -- see mdb_group_keys."""
        torch = _torch()
        F, N = int(keys.shape[0]), int(keys.shape[1])
        cap = int(table[0].numel())
        codes = torch.empty((F, N), dtype=torch.int16, device=self.device)
        order = torch.empty(F * N, dtype=torch.int32, device=self.device) if want_order else None
        seg = torch.empty(cap + 1, dtype=torch.int64, device=self.device) if want_order else None
        if keep is not None:
            keep = keep.to(self.device, dtype=torch.uint8).contiguous()
        rc = self.L.mdb_group_keys(self.ctx, F, N, _ptr(keys), _ptr(keep), int(step0), cap, _ptr(table[0]), _ptr(table[1]),
                                   _ptr(table[2]), _ptr(codes), _ptr(order), _ptr(seg), self.stream())
        _lib.check(rc, "mdb_group_keys")
        return codes, order, seg

    def type_maxcount(self, codes, counts, maxc, targets=None):
        """maxc [cap, Engine.MAXT] int32 (in/out): element-wise maximum of the compositions per key slot."""
        T = int(counts.shape[0])
        rc = self.L.mdb_type_maxcount(self.ctx, _ptr(targets), T, _ptr(codes), _ptr(counts), _ptr(maxc), self.stream())
        _lib.check(rc, "mdb_type_maxcount")

    def minmax_accumulate(self, X, mn, mx):
        rows, D = X.shape
        rc = self.L.mdb_minmax_accumulate(self.ctx, int(rows), int(D), _ptr(X), int(X.stride(0)), _ptr(mn), _ptr(mx),
                                          self.stream())
        _lib.check(rc, "mdb_minmax_accumulate")

    def gather_rows(self, X, src, dst, out):
        """out[dst[r]] = X[src[r]] (src / dst int64 device tensors)."""
        n = int(src.numel())
        rc = self.L.mdb_gather_rows(self.ctx, n, int(X.shape[1]), _ptr(X), int(X.stride(0)), _ptr(src), _ptr(dst), _ptr(out),
                                    int(out.stride(0)), self.stream())
        _lib.check(rc, "mdb_gather_rows")

    def label_hist(self, labels, hist):
        """hist [K] int32 += member counts of labels [rows] int32."""
        rc = self.L.mdb_label_hist(self.ctx, int(labels.numel()), int(hist.numel()), _ptr(labels), _ptr(hist), self.stream())
        _lib.check(rc, "mdb_label_hist")

    def pick_rows(self, labels, picks):
        """picks [n,4] int64 = (first row, end row, cluster, p) -> row of the p-th member of the cluster in the
        range (ascending), -1 if there is none."""
        torch = _torch()
        picks = torch.as_tensor(picks, dtype=torch.int64, device=self.device).contiguous()
        out = torch.empty(int(picks.shape[0]), dtype=torch.int64, device=self.device)
        rc = self.L.mdb_pick_rows(self.ctx, _ptr(labels), int(picks.shape[0]), _ptr(picks), _ptr(out), self.stream())
        _lib.check(rc, "mdb_pick_rows")
        return out

    def components(self, ell):
        torch = _torch()
        F, N, _ = ell.shape
        labels = torch.empty((F, N), dtype=torch.int32, device=self.device)
        _lib.check(self.L.mdb_components(self.ctx, int(F), int(N), _ptr(ell), _ptr(labels), 0, self.stream()),
                   "mdb_components")
        return labels

    def dps(self, rowptr, col):
        """Exact ``dps`` output (dps.pyx:17-46) for one frame: (mol_atoms, mol_ptr) device tensors."""
        torch = _torch()
        N = int(rowptr.numel()) - 1
        mol_atoms = torch.empty(max(N, 1), dtype=torch.int32, device=self.device)
        mol_ptr = torch.empty(N + 1, dtype=torch.int32, device=self.device)
        nm = C.c_int(0)
        rc = self.L.mdb_dps(self.ctx, N, _ptr(rowptr), _ptr(col), _ptr(mol_atoms), _ptr(mol_ptr), C.byref(nm),
                            self.stream())
        _lib.check(rc, "mdb_dps")
        return mol_atoms[:N], mol_ptr[: nm.value + 1]

    def compositions(self, pos, cell, types, cutoff, targets=None, periodic=True, maxcount=None, members_cap=0):
        """Element composition of every target's cutoff sphere (datasetbuilder.py:312-322).
        Returns (counts [T, ntypes] int32 device tensor, maxcount numpy [ntypes] running max); with
        members_cap > 0 also the membership lists [T, members_cap] (int32, -1 padded, any order) that
        `descriptors(members=...)` can use instead of scanning the cells again."""
        torch = _torch()
        pos = self.to_device_pos(pos)
        types = self.to_device_types(types)
        F, N = int(pos.shape[0]), int(pos.shape[1])
        cells = self._cells(cell, F) if periodic else None
        nt = len(self.atomname)
        if targets is not None:
            targets = torch.as_tensor(targets, dtype=torch.int32, device=self.device).contiguous()
            T = int(targets.numel())
        else:
            T = F * N
        counts = torch.zeros((T, nt), dtype=torch.int32, device=self.device)
        mc = np.zeros(nt, dtype=np.int32) if maxcount is None else np.ascontiguousarray(maxcount, dtype=np.int32).copy()
        if members_cap > 0:
            members = torch.empty((T, int(members_cap)), dtype=torch.int32, device=self.device)
            rc = self.L.mdb_compositions_members(self.ctx, F, N, _ptr(pos), _np_ptr(cells), _ptr(types), int(bool(periodic)),
                                                 float(cutoff), _ptr(targets), T, _ptr(counts), _np_ptr(mc),
                                                 int(members_cap), _ptr(members), self.stream())
            _lib.check(rc, "mdb_compositions_members")
            return counts, mc, members
        rc = self.L.mdb_compositions(self.ctx, F, N, _ptr(pos), _np_ptr(cells), _ptr(types), int(bool(periodic)),
                                     float(cutoff), _ptr(targets), T, _ptr(counts), _np_ptr(mc), self.stream())
        _lib.check(rc, "mdb_compositions")
        return counts, mc

    def sphere_members(self, pos, cell, types, cutoff, targets, periodic=True, cap=256):
        """Ascending member lists of the cutoff spheres of `targets` (datasetbuilder.py:554-557)."""
        torch = _torch()
        pos = self.to_device_pos(pos)
        types = self.to_device_types(types)
        F, N = int(pos.shape[0]), int(pos.shape[1])
        cells = self._cells(cell, F) if periodic else None
        targets = torch.as_tensor(targets, dtype=torch.int32, device=self.device).contiguous()
        T = int(targets.numel())
        while True:
            members = torch.empty((T, cap), dtype=torch.int32, device=self.device)
            nmem = torch.empty(T, dtype=torch.int32, device=self.device)
            rc = self.L.mdb_sphere_members(self.ctx, F, N, _ptr(pos), _np_ptr(cells), _ptr(types), int(bool(periodic)),
                                           float(cutoff), _ptr(targets), T, int(cap), _ptr(members), _ptr(nmem),
                                           self.stream())
            _lib.check(rc, "mdb_sphere_members")
            n = nmem.cpu().numpy()
            if T == 0 or n.max() <= cap:
                break
            if cap >= 1024:
                raise RuntimeError(f"a cutoff sphere holds {int(n.max())} atoms; the member kernel lists at most 1024 "
                                   "(reduce the cutoff)")
            cap = min(1024, 2 * cap)
        m = members.cpu().numpy()
        return [m[k, : min(n[k], cap)] for k in range(T)]

    def descriptors(self, pos, cell, types, cutoff, counts, maxcount, targets=None, periodic=True, want_X=True,
                    want_eig=False, eigcap=160, members=None):
        """Coulomb-matrix spectrum + padded sorted feature rows (datasetbuilder.py:236-349).
        members = the lists from `compositions(members_cap=...)` for the same targets: the cluster is
        then gathered from them (no cell list, `cutoff` unused)."""
        torch = _torch()
        pos = self.to_device_pos(pos)
        types = self.to_device_types(types)
        F, N = int(pos.shape[0]), int(pos.shape[1])
        cells = self._cells(cell, F) if periodic else None
        if targets is not None:
            targets = torch.as_tensor(targets, dtype=torch.int32, device=self.device).contiguous()
            T = int(targets.numel())
        else:
            T = F * N
        mc = np.ascontiguousarray(maxcount, dtype=np.int32)
        D = int(mc.sum())
        X = torch.empty((T, D), dtype=torch.float64, device=self.device) if want_X else None
        eig = torch.empty((T, eigcap), dtype=torch.float64, device=self.device) if want_eig else None
        n = torch.empty((T,), dtype=torch.int32, device=self.device)
        if members is not None:
            rc = self.L.mdb_descriptors_members(self.ctx, F, N, _ptr(pos), _np_ptr(cells), _ptr(types), int(bool(periodic)),
                                                _ptr(targets), T, _ptr(counts), _np_ptr(mc), _ptr(members),
                                                int(members.shape[1]), _ptr(X), _ptr(eig), int(eigcap), _ptr(n),
                                                self.stream())
            _lib.check(rc, "mdb_descriptors_members")
            return {"X": X, "eig": eig, "n": n}
        rc = self.L.mdb_descriptors(self.ctx, F, N, _ptr(pos), _np_ptr(cells), _ptr(types), int(bool(periodic)),
                                    float(cutoff), _ptr(targets), T, _ptr(counts), _np_ptr(mc), _ptr(X), _ptr(eig),
                                    int(eigcap), _ptr(n), self.stream())
        _lib.check(rc, "mdb_descriptors")
        return {"X": X, "eig": eig, "n": n}

    def descriptors_compact(self, pos, cell, types, counts, maxcount, targets, members, total_n, periodic=True, want_X=True):
        """`descriptors(members=...)` that also keeps the spectra compactly: returns (X or None, eig_flat
        [total_n] float64, eig_off [T + 1] int32); total_n = counts.sum().  `feature_rows_compact` rebuilds the
        same rows from them later."""
        torch = _torch()
        pos = self.to_device_pos(pos)
        types = self.to_device_types(types)
        F, N = int(pos.shape[0]), int(pos.shape[1])
        cells = self._cells(cell, F) if periodic else None
        targets = torch.as_tensor(targets, dtype=torch.int32, device=self.device).contiguous()
        T = int(targets.numel())
        mc = np.ascontiguousarray(maxcount, dtype=np.int32)
        X = torch.empty((T, int(mc.sum())), dtype=torch.float64, device=self.device) if want_X else None
        eig = torch.empty(max(int(total_n), 1), dtype=torch.float64, device=self.device)
        off = torch.empty(T + 1, dtype=torch.int32, device=self.device)
        n = torch.empty(T, dtype=torch.int32, device=self.device)
        rc = self.L.mdb_descriptors_members_compact(self.ctx, F, N, _ptr(pos), _np_ptr(cells), _ptr(types), int(bool(periodic)),
                                                    _ptr(targets), T, _ptr(counts), _np_ptr(mc), _ptr(members),
                                                    int(members.shape[1]), _ptr(X), _ptr(eig), _ptr(off), _ptr(n), self.stream())
        _lib.check(rc, "mdb_descriptors_members_compact")
        return X, eig, off

    def feature_rows_compact(self, eig, off, counts, maxcount):
        """Padded sorted rows [T, D] from compact spectra (see descriptors_compact)."""
        torch = _torch()
        mc = np.ascontiguousarray(maxcount, dtype=np.int32)
        T = int(off.numel()) - 1
        X = torch.empty((T, int(mc.sum())), dtype=torch.float64, device=self.device)
        rc = self.L.mdb_feature_rows_compact(self.ctx, T, _ptr(eig), _ptr(off), _ptr(counts), _np_ptr(mc), _ptr(X), self.stream())
        _lib.check(rc, "mdb_feature_rows_compact")
        return X

    def feature_rows(self, eig, n, counts, maxcount):
        """Padded sorted feature rows [T, D] from stored spectra (eig [T, eigcap], n [T]) and compositions once
        the final per-element maxima are known (closed form of datasetbuilder.py:236-276)."""
        torch = _torch()
        mc = np.ascontiguousarray(maxcount, dtype=np.int32)
        T = int(eig.shape[0])
        X = torch.empty((T, int(mc.sum())), dtype=torch.float64, device=self.device)
        rc = self.L.mdb_feature_rows(self.ctx, T, _ptr(eig), int(eig.shape[1]), _ptr(n), _ptr(counts), _np_ptr(mc), _ptr(X),
                                     self.stream())
        _lib.check(rc, "mdb_feature_rows")
        return X

    # ------------------------------------------------------------------ clustering
    def minmax_fit(self, X):
        torch = _torch()
        rows, D = X.shape
        mn = torch.empty(D, dtype=torch.float64, device=self.device)
        mx = torch.empty(D, dtype=torch.float64, device=self.device)
        _lib.check(self.L.mdb_minmax_fit(self.ctx, int(rows), int(D), _ptr(X), int(X.stride(0)), _ptr(mn), _ptr(mx),
                                         self.stream()), "mdb_minmax_fit")
        return mn, mx

    def minmax_transform(self, X, data_min, data_max):
        """In-place MinMaxScaler.transform; data_min / data_max are host arrays [D]."""
        rows, D = X.shape
        mn = np.ascontiguousarray(data_min, dtype=np.float64)
        mx = np.ascontiguousarray(data_max, dtype=np.float64)
        _lib.check(self.L.mdb_minmax_transform(self.ctx, int(rows), int(D), _ptr(X), int(X.stride(0)), _np_ptr(mn),
                                               _np_ptr(mx), self.stream()), "mdb_minmax_transform")
        return X

    def kmeans_assign(self, X, centers, rowidx=None, weights=None, want_inertia=False, want_mindist=False):
        torch = _torch()
        D = int(X.shape[1])
        K = int(centers.shape[0])
        rows = int(rowidx.numel()) if rowidx is not None else int(X.shape[0])
        labels = torch.empty(rows, dtype=torch.int32, device=self.device)
        mind = torch.empty(rows, dtype=torch.float64, device=self.device) if want_mindist else None
        inertia = C.c_double(0.0)
        rc = self.L.mdb_kmeans_assign(self.ctx, rows, D, K, _ptr(X), int(X.stride(0)), _ptr(rowidx), _ptr(centers),
                                      _ptr(labels), _ptr(mind), _ptr(weights),
                                      C.byref(inertia) if want_inertia else None, self.stream())
        _lib.check(rc, "mdb_kmeans_assign")
        return labels, (inertia.value if want_inertia else None), mind

    def kmeans_assign_tc(self, X, centers, rowidx=None, diag=False):
        """Labels via the tensor-core screening pass + float64 re-check of near-ties.
        Returns (labels, number of rows re-checked in float64)."""
        torch = _torch()
        D = int(X.shape[1])
        K = int(centers.shape[0])
        rows = int(rowidx.numel()) if rowidx is not None else int(X.shape[0])
        labels = torch.empty(rows, dtype=torch.int32, device=self.device)
        nre = C.c_longlong(0)
        dg = torch.empty(4 * rows + 1, dtype=torch.float32, device=self.device) if diag else None
        rc = self.L.mdb_kmeans_assign_tc(self.ctx, rows, D, K, _ptr(X), int(X.stride(0)), _ptr(rowidx), _ptr(centers),
                                         _ptr(labels), C.byref(nre), _ptr(dg), self.stream())
        _lib.check(rc, "mdb_kmeans_assign_tc")
        if diag:
            return labels, int(nre.value), dg
        return labels, int(nre.value)

    def kmeans_inertia(self, X, centers, labels, rowidx=None, weights=None):
        rows = int(labels.numel())
        v = C.c_double(0.0)
        rc = self.L.mdb_kmeans_inertia(self.ctx, rows, int(X.shape[1]), _ptr(X), int(X.stride(0)), _ptr(rowidx),
                                       _ptr(centers), _ptr(labels), _ptr(weights), C.byref(v), self.stream())
        _lib.check(rc, "mdb_kmeans_inertia")
        return v.value

    def kmeans_partial_sums(self, X, labels, K, rowidx=None, weights=None, out=None):
        torch = _torch()
        D = int(X.shape[1])
        B = int(labels.numel())
        if out is None:
            sums = torch.empty((K, D), dtype=torch.float64, device=self.device)
            wsum = torch.empty(K, dtype=torch.float64, device=self.device)
        else:
            sums, wsum = out
        rc = self.L.mdb_kmeans_partial_sums(self.ctx, B, D, int(K), _ptr(X), int(X.stride(0)), _ptr(rowidx),
                                            _ptr(weights), _ptr(labels), _ptr(sums), _ptr(wsum), self.stream())
        _lib.check(rc, "mdb_kmeans_partial_sums")
        return sums, wsum

    def kmeans_update_exact(self, X, labels, centers, weight_sums, rowidx=None, weights=None):
        """Mini-batch M-step (<= 4096 rows) in scikit-learn's operation order, in place."""
        rc = self.L.mdb_kmeans_update_exact(self.ctx, int(labels.numel()), int(X.shape[1]), _ptr(X), int(X.stride(0)),
                                            _ptr(rowidx), _ptr(weights), _ptr(labels), _ptr(centers), _ptr(weight_sums),
                                            self.stream())
        _lib.check(rc, "mdb_kmeans_update_exact")

    def kmeans_minibatch_step(self, X, idx, centers, weight_sums):
        """One fused mini-batch step: idx = host int32 array of row indices.  Returns (batch inertia before the
        update, number of centres whose weight is still zero after it); centers / weight_sums change in place."""
        idx = np.ascontiguousarray(idx, dtype=np.int32)
        inertia = C.c_double(0.0)
        nzero = C.c_int(0)
        rc = self.L.mdb_kmeans_minibatch_step(self.ctx, int(idx.size), int(X.shape[1]), int(centers.shape[0]), _ptr(X),
                                              int(X.stride(0)), _np_ptr(idx), _ptr(centers), _ptr(weight_sums),
                                              C.byref(inertia), C.byref(nzero), self.stream())
        _lib.check(rc, "mdb_kmeans_minibatch_step")
        return inertia.value, nzero.value

    def kpp_step(self, X, cand, closest, dist_out, pot_out, rowidx=None, weights=None):
        """One k-means++ seeding step: distances of the candidate rows to every row (clipped
        against `closest`) and the weighted potential of each candidate."""
        n = int(rowidx.numel()) if rowidx is not None else int(X.shape[0])
        rc = self.L.mdb_kpp_step(self.ctx, n, int(X.shape[1]), _ptr(X), int(X.stride(0)), _ptr(rowidx), _ptr(cand),
                                 int(cand.numel()), _ptr(closest), _ptr(weights), _ptr(dist_out), _ptr(pot_out),
                                 self.stream())
        _lib.check(rc, "mdb_kpp_step")

    def kmeans_plusplus(self, X, K, rand, first_center, rowidx=None):
        """k-means++ seeding on the device (sklearn _kmeans_plusplus).  Single run: `rand` = the
        (K-1, n_local_trials) uniform variates in drawing order (host), first_center an int, rowidx [n] or
        None; returns the K chosen positions (int32 device tensor).  Batch of independent runs: rand
        [B, K-1, T], first_center [B], rowidx [B, n] or None; returns [B, K]."""
        torch = _torch()
        rand = np.asarray(rand, dtype=np.float64)
        batched = np.ndim(first_center) > 0
        B = len(first_center) if batched else 1
        T = int(rand.shape[-1])
        rand = np.ascontiguousarray(rand.reshape(B, -1))
        first = np.ascontiguousarray(np.atleast_1d(first_center), dtype=np.int32)
        if rowidx is not None:
            rowidx = rowidx.contiguous()
            n = int(rowidx.shape[-1])
        else:
            n = int(X.shape[0])
        idx = torch.empty((B, K), dtype=torch.int32, device=self.device)
        rc = self.L.mdb_kmeans_plusplus(self.ctx, B, n, int(X.shape[1]), _ptr(X), int(X.stride(0)), _ptr(rowidx), int(K), T,
                                        _np_ptr(rand), _np_ptr(first), _ptr(idx), self.stream())
        _lib.check(rc, "mdb_kmeans_plusplus")
        return idx if batched else idx[0]

    def kmeans_apply_update(self, centers, weight_sums, sums, wsum):
        K, D = centers.shape
        rc = self.L.mdb_kmeans_apply_update(self.ctx, int(K), int(D), _ptr(centers), _ptr(weight_sums), _ptr(sums),
                                            _ptr(wsum), self.stream())
        _lib.check(rc, "mdb_kmeans_apply_update")

    def analyze_host(self, pos, cell, types, periodic=True, want_keys=True, want_labels=True, out=None,
                     ell_width=0, key_codes=False, edges=False):
        """Host-buffer path (e2e): numpy / pinned tensors in, numpy out; H2D and D2H copies
        are pipelined inside the library.  key_codes=True returns the bond-type keys as one byte per
        atom ("keycode") plus the dictionary "keytable" (256 x uint64, unused = all ones) instead of
        "keys".  edges=True (implies key codes) returns the bonds as a list, each once: "edges" [E, 2] with
        i < j and "edge_ptr" [F + 1] (frame f owns edges[edge_ptr[f]:edge_ptr[f + 1]]) instead of "ell"."""
        torch = _torch()
        if isinstance(pos, torch.Tensor):
            pos_np = pos.numpy()
        else:
            pos_np = np.ascontiguousarray(pos, dtype=np.float64)
        if pos_np.ndim == 2:
            pos_np = pos_np[None]
        F, N = pos_np.shape[0], pos_np.shape[1]
        types_np = np.ascontiguousarray(np.asarray(types).astype(np.uint8))
        cells = self._cells(cell, F) if periodic else None
        out = out or {}
        ell = out.get("ell")
        if ell is None and not edges:
            ell = np.empty((F, N, ell_width or self.max_bonds), dtype=np.int32)
        labels = out.get("labels") if want_labels else None
        if want_labels and labels is None:
            labels = np.empty((F, N), dtype=np.int32)

        def p(a):
            if a is None:
                return C.c_void_p(0)
            return C.c_void_p(a.data_ptr()) if isinstance(a, torch.Tensor) else C.c_void_p(a.ctypes.data)

        if edges:
            ebuf = out.get("edges")
            if ebuf is None:
                ebuf = np.empty((F * N * 2, 2), dtype=np.int32)   # room for an average degree of 4
            eptr = out.get("edge_ptr")
            if eptr is None:
                eptr = np.empty(F + 1, dtype=np.int64)
            code = out.get("keycode") if want_keys else None
            if want_keys and code is None:
                code = np.empty((F, N), dtype=np.uint8)
            table = out.get("keytable") if want_keys else None
            if want_keys and table is None:
                table = np.empty(256, dtype=np.uint64)
            ecap = int(ebuf.shape[0]) if not isinstance(ebuf, torch.Tensor) else int(ebuf.shape[0])
            rc = self.L.mdb_analyze_frames_host_edges(self.ctx, int(F), int(N), p(pos_np), _np_ptr(cells), p(types_np),
                                                      int(bool(periodic)), p(ebuf), ecap, p(eptr), p(code), p(table),
                                                      p(labels))
            _lib.check(rc, "mdb_analyze_frames_host_edges")
            return {"edges": ebuf, "edge_ptr": eptr, "keycode": code, "keytable": table, "labels": labels}
        if key_codes:
            code = out.get("keycode")
            if code is None:
                code = np.empty((F, N), dtype=np.uint8)
            table = out.get("keytable")
            if table is None:
                table = np.empty(256, dtype=np.uint64)
            rc = self.L.mdb_analyze_frames_host_coded(self.ctx, int(F), int(N), p(pos_np), _np_ptr(cells), p(types_np),
                                                      int(bool(periodic)), p(ell), int(ell_width), p(code), p(table),
                                                      p(labels))
            _lib.check(rc, "mdb_analyze_frames_host_coded")
            return {"ell": ell, "keycode": code, "keytable": table, "labels": labels}
        keys = out.get("keys") if want_keys else None
        if want_keys and keys is None:
            keys = np.empty((F, N), dtype=np.uint64)
        rc = self.L.mdb_analyze_frames_host(self.ctx, int(F), int(N), p(pos_np), _np_ptr(cells), p(types_np),
                                            int(bool(periodic)), p(ell), int(ell_width), p(keys), p(labels))
        _lib.check(rc, "mdb_analyze_frames_host")
        return {"ell": ell, "keys": keys, "labels": labels}


_ENGINES = {}


def get_engine(device: int = 0, atomname=("C", "H", "O")) -> Engine:
    """Process-wide engine per device (re-targeted when the element list changes;
    atomname=None keeps whatever table is loaded)."""
    e = _ENGINES.get(device)
    if e is None:
        e = Engine(device, atomname if atomname is not None else ("C", "H", "O"))
        _ENGINES[device] = e
    elif atomname is not None and [str(a) for a in atomname] != e.atomname:
        e.set_elements(atomname)
    return e
