"""LAMMPS text readers backed by the C++ parser in libmdb_b200.so (csrc/textio.cpp).

Mirrors what the reference extracts from its input files: ``_readN`` (detect.py:56-88,
151-194) at open time, then frames sliced as fixed groups of ``steplinenum`` lines with a
stride (datasetbuilder.py:616-638).
"""

from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .utils import must_be_list

KIND_DUMP = 0
KIND_BOND = 1


class TextReader:
    def __init__(self, kind: int, filename, nthreads: int = 0):
        self.L = _lib.lib()
        self.kind = kind
        self.paths = [str(p) for p in must_be_list(filename)]
        arr = (C.c_char_p * len(self.paths))(*[p.encode() for p in self.paths])
        h = C.c_void_p()
        rc = self.L.mdb_reader_open(kind, arr, len(self.paths), int(nthreads), C.byref(h))
        if rc != 0:
            raise RuntimeError(_lib.last_error())
        self.h = h
        n = C.c_int()
        sl = C.c_long()
        cols = (C.c_int * 5)()
        _lib.check(self.L.mdb_reader_info(self.h, C.byref(n), C.byref(sl), cols), "mdb_reader_info")
        self.natoms = n.value
        self.steplinenum = sl.value
        self.cols = list(cols)
        t = np.zeros(self.natoms, dtype=np.int32)
        _lib.check(self.L.mdb_reader_types(self.h, C.c_void_p(t.ctypes.data)), "mdb_reader_types")
        self.atomtype = t  # LAMMPS types (1-based) by id-1

    def close(self):
        if getattr(self, "h", None):
            self.L.mdb_reader_close(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def rewind(self):
        _lib.check(self.L.mdb_reader_rewind(self.h), "mdb_reader_rewind")

    def read(self, max_frames: int, stride: int = 1, width: int = 8, want_bo: bool = True, pinned=None):
        """Next batch of kept frames.  dump -> (pos [F,N,3], cell [F,9]); bonds ->
        (nbr [F,N,width] int32, bo [F,N,width] float64 or None).  F == 0 at end of input."""
        N = self.natoms
        nread = C.c_int(0)
        if self.kind == KIND_DUMP:
            pos = np.empty((max_frames, N, 3), dtype=np.float64) if pinned is None else pinned
            cell = np.empty((max_frames, 9), dtype=np.float64)
            rc = self.L.mdb_reader_next(self.h, int(stride), int(max_frames), C.c_void_p(pos.ctypes.data),
                                        C.c_void_p(cell.ctypes.data), None, None, 0, C.byref(nread))
            if rc != 0:
                raise RuntimeError(_lib.last_error())
            return pos[: nread.value], cell[: nread.value]
        nbr = np.empty((max_frames, N, width), dtype=np.int32)
        bo = np.empty((max_frames, N, width), dtype=np.float64) if want_bo else None
        rc = self.L.mdb_reader_next(self.h, int(stride), int(max_frames), None, None, C.c_void_p(nbr.ctypes.data),
                                    C.c_void_p(bo.ctypes.data) if bo is not None else None, int(width),
                                    C.byref(nread))
        if rc != 0:
            raise RuntimeError(_lib.last_error())
        return nbr[: nread.value], (bo[: nread.value] if bo is not None else None)

    def skip(self, max_frames: int, stride: int = 1) -> int:
        """Step over the next batch of kept frames without parsing; returns how many."""
        nread = C.c_int(0)
        rc = self.L.mdb_reader_next(self.h, int(stride), int(max_frames), None, None, None, None, 0, C.byref(nread))
        if rc != 0:
            raise RuntimeError(_lib.last_error())
        return nread.value

    def parse_lines(self, lines, width: int = 8, want_bo: bool = True):
        """Parse ONE frame given as an iterable of text lines (None entries are skipped,
        like the `if line:` guards at detect.py:109,139,303)."""
        buf = "".join(l if l.endswith("\n") else l + "\n" for l in lines if l).encode()
        N = self.natoms
        if self.kind == KIND_DUMP:
            pos = np.empty((N, 3), dtype=np.float64)
            cell = np.empty(9, dtype=np.float64)
            rc = self.L.mdb_reader_parse_buffer(self.h, buf, len(buf), C.c_void_p(pos.ctypes.data),
                                                C.c_void_p(cell.ctypes.data), None, None, 0)
            if rc != 0:
                raise RuntimeError(_lib.last_error())
            return pos, cell.reshape(3, 3)
        nbr = np.empty((N, width), dtype=np.int32)
        bo = np.empty((N, width), dtype=np.float64) if want_bo else None
        rc = self.L.mdb_reader_parse_buffer(self.h, buf, len(buf), None, None, C.c_void_p(nbr.ctypes.data),
                                            C.c_void_p(bo.ctypes.data) if bo is not None else None, int(width))
        if rc != 0:
            raise RuntimeError(_lib.last_error())
        return nbr, bo
