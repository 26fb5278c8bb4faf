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
        self._drop_prefetch()
        if getattr(self, "h", None):
            self.L.mdb_reader_close(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def rewind(self):
        self._drop_prefetch()
        _lib.check(self.L.mdb_reader_rewind(self.h), "mdb_reader_rewind")

    def read(self, max_frames: int, stride: int = 1, width: int = 8, want_bo: bool = True, pinned=None):
        """Next batch of kept frames.  dump -> (pos [F,N,3], cell [F,9]); bonds ->
        (nbr [F,N,width] int32, bo [F,N,width] float64 or None).  F == 0 at end of input."""
        if getattr(self, "_pf_thread", None) is not None:
            raise RuntimeError("a prefetched device batch is pending (read_device(prefetch=True)); rewind() drops it")
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

    def _stage_batch(self, slot: int, max_frames: int, stride: int):
        """Host half of read_device: slice the next kept frames into page-locked buffer `slot`."""
        import torch

        if getattr(self, "_frame_bytes", None) is None:
            # staging sized from the first frame of the first file (+25 %), at least 1 MiB per frame
            self._frame_bytes = max(1 << 20, int(1.25 * self._first_frame_bytes()))
            self._stages = [None, None]
        fb = np.empty(max_frames, dtype=np.int64)
        ab = np.empty(max_frames, dtype=np.int64)
        fe = np.empty(max_frames, dtype=np.int64)
        cell = np.empty((max_frames, 9), dtype=np.float64)
        nread = C.c_int(0)
        while True:
            cap = int(self._frame_bytes) * int(max_frames)
            if self._stages[slot] is None or self._stages[slot].numel() < cap:
                self._stages[slot] = torch.empty(cap, dtype=torch.uint8, pin_memory=True)
            st = self._stages[slot]
            rc = self.L.mdb_reader_next_raw(self.h, int(stride), int(max_frames), C.c_void_p(st.data_ptr()),
                                            C.c_size_t(st.numel()), C.c_void_p(fb.ctypes.data),
                                            C.c_void_p(ab.ctypes.data), C.c_void_p(fe.ctypes.data),
                                            C.c_void_p(cell.ctypes.data), C.byref(nread))
            if rc == 0:
                break
            msg = _lib.last_error()
            if "does not fit" not in msg or st.numel() > (1 << 33):
                return {"error": msg}
            self._frame_bytes *= 2  # a frame larger than the estimate (the reader has not moved): grow
        return {"F": nread.value, "fb": fb, "ab": ab, "fe": fe, "cell": cell, "stage": st, "args": (max_frames, stride)}

    def _drop_prefetch(self):
        t = getattr(self, "_pf_thread", None)
        if t is not None:
            t.join()
            self._pf_thread = None
            self._pf_result = None

    def read_device(self, max_frames: int, engine, stride: int = 1, prefetch: bool = False, width: int = 8,
                    want_bo: bool = True):
        """Next batch of kept frames, parsed ON THE GPU: the raw text goes through a page-locked
        staging buffer to the device, `mdb_parse_dump_text` / `mdb_parse_bonds_text` turn the atom
        lines into arrays.  dump -> (pos [F,N,3] float64 device tensor, cell [F,9] numpy); bonds ->
        (nbr [F,N,width] int32 device tensor, bo [F,N,width] float64 device tensor or None).
        F == 0 at end of input.
        Frames the device parser hands back (numbers outside its exact fast path, malformed lines)
        are parsed by the host parser, which also owns the error messages.
        prefetch=True promises that the next call on this reader is the same read_device call: the
        host then stages that batch on a background thread while the GPU (and the caller) work on
        this one."""
        import threading

        import torch

        N = self.natoms
        slot = getattr(self, "_slot", 0)
        batch = None
        t = getattr(self, "_pf_thread", None)
        if t is not None:
            t.join()
            batch = self._pf_result
            self._pf_thread = None
            self._pf_result = None
            if "error" not in batch and batch["args"] != (max_frames, stride):
                raise RuntimeError("read_device: a prefetched batch is pending with other arguments")
        if batch is None:
            batch = self._stage_batch(slot, max_frames, stride)
        if "error" in batch:
            raise RuntimeError(batch["error"])
        self._slot = slot ^ 1
        F = batch["F"]
        dev = engine.device
        if F == 0:
            if self.kind == KIND_BOND:
                return (torch.empty((0, N, width), dtype=torch.int32, device=dev),
                        torch.empty((0, N, width), dtype=torch.float64, device=dev) if want_bo else None)
            return torch.empty((0, N, 3), dtype=torch.float64, device=dev), batch["cell"][:0]
        if prefetch:
            def work(s=self._slot):
                self._pf_result = self._stage_batch(s, max_frames, stride)

            self._pf_thread = threading.Thread(target=work, daemon=True)
            self._pf_thread.start()
        fb, ab, fe, cell, stage = batch["fb"], batch["ab"], batch["fe"], batch["cell"], batch["stage"]
        total = int(fe[F - 1])
        d_text = torch.empty(total + 32, dtype=torch.uint8, device=dev)
        d_text[:total].copy_(stage[:total], non_blocking=True)
        d_ab = torch.from_numpy(ab[:F]).to(dev)
        d_fe = torch.from_numpy(fe[:F]).to(dev)
        status = torch.empty(F, dtype=torch.int32, device=dev)
        count = torch.empty(F, dtype=torch.int32, device=dev)
        if self.kind == KIND_BOND:
            nbr = torch.empty((F, N, width), dtype=torch.int32, device=dev)
            bo = torch.empty((F, N, width), dtype=torch.float64, device=dev) if want_bo else None
            rc = self.L.mdb_parse_bonds_text(engine.ctx, C.c_void_p(d_text.data_ptr()), F, N, C.c_void_p(d_ab.data_ptr()),
                                             C.c_void_p(d_fe.data_ptr()), C.c_longlong(int((fe[:F] - fb[:F]).max())),
                                             int(width), C.c_void_p(nbr.data_ptr()),
                                             C.c_void_p(bo.data_ptr()) if bo is not None else None,
                                             C.c_void_p(status.data_ptr()), C.c_void_p(count.data_ptr()), engine.stream())
            _lib.check(rc, "mdb_parse_bonds_text")
            st = status.cpu().numpy()  # synchronises: this staging buffer is free again after this
            self.device_fallback_frames = getattr(self, "device_fallback_frames", 0) + int((st != 0).sum())
            for k in np.nonzero(st)[0]:
                hn = np.empty((N, width), dtype=np.int32)
                hb = np.empty((N, width), dtype=np.float64) if want_bo else None
                base = stage.data_ptr() + int(fb[k])
                rc = self.L.mdb_reader_parse_buffer(self.h, C.cast(C.c_void_p(base), C.c_char_p), int(fe[k] - fb[k]), None,
                                                    None, C.c_void_p(hn.ctypes.data),
                                                    C.c_void_p(hb.ctypes.data) if hb is not None else None, int(width))
                if rc != 0:
                    msg = _lib.last_error()
                    self._drop_prefetch()
                    raise RuntimeError(msg)
                nbr[k].copy_(torch.from_numpy(hn))
                if bo is not None:
                    bo[k].copy_(torch.from_numpy(hb))
            return nbr, bo
        if getattr(self, "_d_types", None) is None or self._d_types.device != dev:
            self._d_types = torch.from_numpy(self.atomtype.astype(np.int32)).to(dev)
        pos = torch.empty((F, N, 3), dtype=torch.float64, device=dev)
        cols = (C.c_int * 5)(*self.cols)
        rc = self.L.mdb_parse_dump_text(engine.ctx, C.c_void_p(d_text.data_ptr()), F, N, C.c_void_p(d_ab.data_ptr()),
                                        C.c_void_p(d_fe.data_ptr()), C.c_longlong(int((fe[:F] - fb[:F]).max())), cols,
                                        C.c_void_p(self._d_types.data_ptr()), C.c_void_p(pos.data_ptr()),
                                        C.c_void_p(status.data_ptr()), C.c_void_p(count.data_ptr()), engine.stream())
        _lib.check(rc, "mdb_parse_dump_text")
        st = status.cpu().numpy()  # synchronises: this staging buffer is free again after this
        self.device_fallback_frames = getattr(self, "device_fallback_frames", 0) + int((st != 0).sum())
        for k in np.nonzero(st)[0]:
            hp = np.empty((N, 3), dtype=np.float64)
            hc = np.empty(9, dtype=np.float64)
            base = stage.data_ptr() + int(fb[k])
            rc = self.L.mdb_reader_parse_buffer(self.h, C.cast(C.c_void_p(base), C.c_char_p), int(fe[k] - fb[k]),
                                                C.c_void_p(hp.ctypes.data), C.c_void_p(hc.ctypes.data), None, None, 0)
            if rc != 0:
                msg = _lib.last_error()
                self._drop_prefetch()
                raise RuntimeError(msg)
            pos[k].copy_(torch.from_numpy(hp))
            cell[k] = hc
        return pos, cell[:F]

    def _first_frame_bytes(self) -> int:
        n = 0
        with open(self.paths[0], "rb") as f:
            for _ in range(int(self.steplinenum)):
                line = f.readline()
                if not line:
                    break
                n += len(line)
        return n

    def skip(self, max_frames: int, stride: int = 1) -> int:
        """Step over the next batch of kept frames without parsing; returns how many."""
        if getattr(self, "_pf_thread", None) is not None:
            raise RuntimeError("a prefetched device batch is pending (read_device(prefetch=True)); rewind() drops it")
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
