"""Streaming selection: bond typing, descriptors, clustering and per-cluster choice of EVERY bond type in
three passes over the frames; no feature matrix is ever held for all rows (SURVEY.md §7 hard part 3) -- the
compact spectra are, in HBM and host memory, as long as they fit, and are recomputed where they do not.

The reference holds, per bond type, the whole feature matrix in RAM (``feedvector``, reference
datasetbuilder.py:238,255-260) and re-reads the trajectory once per type (:243).  At 1e9 atom-frames
neither fits anywhere, so the steps of ``_readtimestepsbond`` / ``_writecoulumbmatrix`` /
``_clusterdatas`` (datasetbuilder.py:185-281, 351-384) are regrouped by what they need to know:

pass A  keys of every atom-frame -> device key dictionary + stable group-by (the per-type ``(step, ids)``
        lists as index arrays), sphere compositions -> ``max_counter`` per type, molecule labels;
pass B  descriptors of all types at once -> running column bounds of the scaler (``MinMaxScaler`` needs
        nothing else), rows of the seeded training pool collected on the fly;
train   ``MiniBatchKMeans`` per type on its scaled pool (all rows when the type has no more than
        ``pool_rows`` of them -- then this is exactly the in-memory fit);
pass C  rows again (from the spectra pass B kept; recomputed only when neither cache tier had room) -> scale ->
        assignment (tcgen05 screening + float64 re-check) -> labels and per-(batch, cluster) member counts;
select  ``default_rng(seed)`` draws one member position per cluster and draw, resolved against the stored
        labels on the device -- the same rows ``choose_per_cluster`` picks from in-memory labels.

Spectra kept from pass B for pass C live in two tiers: HBM while the budget lasts, then ordinary host memory
behind a small page-locked ring (``_HostTier``): device -> ring -> pageable copies trail pass B on a copy stream
and a worker thread, and two prefetchers walk pass C's batch order ahead of the compute stream.  Page-locking the
whole tier is not an option (measured on the B200 box: 2 GiB/s, and kernel launches stall meanwhile).

Multi-GPU: frame batches are dealt round-robin over the ranks; the exchanges are the key vocabulary and
per-batch counts (objects), ``max_counter`` (MAX), scaler bounds (MIN / MAX), the pool rows (SUM of
disjoint contributions), per-batch cluster counts and the chosen rows.  Training runs replicated on the
replicated pool, so results do not depend on the number of ranks.
"""

from __future__ import annotations

import os
import queue
import threading
import time
from dataclasses import dataclass, field
from typing import Callable, Dict, Iterable, List, Optional

import numpy as np

from .kmeans import MiniBatchKMeansB200, _dist

KEY_EMPTY = np.uint64(0xFFFFFFFFFFFFFFFF)


@dataclass
class FrameBatch:
    """One batch of frames on the device.  Bond table: none (dump-only: bonds from coordinates), ELL rows
    (``nbr`` / ``bo`` [F, N, width], what the text parser delivers) or CSR (``rowptr`` [F, N+1], ``col`` /
    ``bo`` [F, colcap])."""
    gidx: int                 # global batch number (defines the global row order across ranks)
    step0: int                # step (frame index after striding) of the first frame
    pos: "object"             # torch [F, N, 3] float64
    cell: np.ndarray          # [F, 9]
    nbr: "object" = None
    bo: "object" = None
    rowptr: "object" = None
    col: "object" = None
    keep: "object" = None     # optional [F, N] uint8 (model-deviation filter)


@dataclass
class _Store:
    gidx: int
    step0: int
    nframes: int
    order: "object"           # int32 [kept] atom-frame indices sorted by key slot
    seg_d: "object"           # int64 [cap + 1] device
    seg: Optional[np.ndarray] = None
    kept: int = 0


@dataclass
class SelectionResult:
    keys: List[int]                         # 64-bit keys in first-seen order
    counts: Dict[int, int]                  # key -> number of atom-frames
    chosen: Dict[int, List[tuple]]          # key -> [(step, atom id 1-based)] in the reference's write order
    maxcount: Dict[int, np.ndarray]         # key -> max_counter of clustered types
    nstep: int = 0
    timings: Dict[str, float] = field(default_factory=dict)
    stats: Dict[str, object] = field(default_factory=dict)


_RING = {}   # (device index, slots, piece) -> page-locked slots, allocated once per process (0.4 s for 768 MB)


def _host_memory_available():
    """Bytes of host memory this process may still take: MemAvailable, capped by the cgroup limit if one is set."""
    avail = None
    try:
        with open("/proc/meminfo") as f:
            for line in f:
                if line.startswith("MemAvailable:"):
                    avail = int(line.split()[1]) * 1024
                    break
    except OSError:
        pass
    for lim, cur in (("/sys/fs/cgroup/memory.max", "/sys/fs/cgroup/memory.current"),
                     ("/sys/fs/cgroup/memory/memory.limit_in_bytes", "/sys/fs/cgroup/memory/memory.usage_in_bytes")):
        try:
            with open(lim) as f:
                v = f.read().strip()
            if v != "max" and int(v) < (1 << 60):
                with open(cur) as f:
                    left = int(v) - int(f.read().strip())
                avail = left if avail is None else min(avail, left)
        except (OSError, ValueError):
            pass
    return avail or 0


class _HostTier:
    """Second tier of the spectra cache: pageable host memory behind a ring of page-locked slots.

    put()  (pass B) hands a batch's device tensors to a worker: device -> slot copies on a copy stream, slot ->
           pageable copies on the worker (first touch of the pages included); the device tensors are released
           when the last piece has landed.  At most `depth` batches wait, so HBM use stays bounded.
    start_fetch() / take()  (pass C): two prefetchers (alternate batches, half of the ring and a copy stream each)
           walk the given batch order, pageable -> slot -> device, and hand over (tensors, ready event) ahead of the
           compute stream.
    The workers' host copies run on a few threads only (`copy_threads`): the main thread launches kernels between
    read-backs and must not wait for a core (measured: pass B 28.7 -> 31.5 s with 16 copy threads on 16 cores).
    """

    def __init__(self, torch, device, budget_bytes, piece=128 << 20, slots=8, depth=2, copy_threads=(4, 6)):
        self.torch, self.device = torch, device
        self.budget = int(budget_bytes)
        self.bytes = 0
        self.piece, self.depth = int(piece), int(depth)
        key = (torch.device(device).index, slots, self.piece)
        if key not in _RING:
            _RING[key] = [torch.empty(self.piece, dtype=torch.uint8, pin_memory=True) for _ in range(slots)]
        self.ring = _RING[key]
        self.cs = torch.cuda.Stream(device)
        self.cs2 = torch.cuda.Stream(device)
        self.copy_threads = copy_threads
        self._threads0 = torch.get_num_threads()
        self._workers, self._outs, self._pos = [], [], {}
        self.wait_put_s, self.wait_take_s, self.spill_busy_s, self.fetch_busy_s = 0.0, 0.0, 0.0, 0.0
        self.store = {}                      # gidx -> {key: [(host uint8 tensor, dtype, shape)]}
        self.index = set()                   # batches handed to put(): what pass C will find here
        self.err = None
        self._q = None
        self._worker = None
        self._out = None

    def __contains__(self, gidx):
        return gidx in self.index

    def __len__(self):
        return len(self.index)

    def room(self, need):
        return self.err is None and self.bytes + need <= self.budget

    # -------------------------------------------------------------- helpers
    def _bytes(self, t):
        return t.reshape(-1).view(self.torch.uint8)

    def _check(self):
        if self.err is not None:
            raise RuntimeError("host tier of the spectra cache failed") from self.err

    def _pump(self, segments, to_host, slots=None, stream=None):
        """Moves (device bytes, host bytes) segments through (part of) the ring, `piece` bytes at a time."""
        torch = self.torch
        free = list(range(len(self.ring))) if slots is None else list(slots)
        stream = self.cs if stream is None else stream
        pending = []                         # (slot, event, host view or None)

        def retire():
            slot, ev, hv = pending.pop(0)
            ev.synchronize()
            if hv is not None:
                hv.copy_(self.ring[slot][: hv.numel()])
            free.append(slot)

        with torch.cuda.stream(stream):
            for dv, hv in segments:
                n = dv.numel()
                for o in range(0, n, self.piece):
                    m = min(self.piece, n - o)
                    if not free:
                        retire()
                    slot = free.pop()
                    sv = self.ring[slot][:m]
                    if to_host:
                        sv.copy_(dv[o:o + m], non_blocking=True)
                    else:
                        sv.copy_(hv[o:o + m])
                        dv[o:o + m].copy_(sv, non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(stream)
                    pending.append((slot, ev, hv[o:o + m] if to_host else None))
            while pending:
                retire()

    # -------------------------------------------------------------- pass B
    def _spill_loop(self):
        torch = self.torch
        try:
            torch.cuda.set_device(self.device)
            torch.set_num_threads(self.copy_threads[0])
            while True:
                job = self._q.get()
                if job is None:
                    return
                gidx, ev, keep = job
                tb = time.perf_counter()
                self.cs.wait_event(ev)
                entry, segments = {}, []
                for k, tensors in keep.items():
                    hs = []
                    for t in tensors:
                        h = torch.empty(t.numel() * t.element_size(), dtype=torch.uint8)
                        hs.append((h, t.dtype, tuple(t.shape)))
                        if t.numel():
                            segments.append((self._bytes(t), h))
                    entry[k] = hs
                self._pump(segments, True)
                self.spill_busy_s += time.perf_counter() - tb
                self.store[gidx] = entry
                del keep, segments, job      # the device tensors go back to the allocator only now
        except BaseException as e:  # noqa: BLE001 -- reported to the main thread by _check()
            self.err = e

    def put(self, gidx, keep, need):
        """keep: {key: (eig, off, counts)} device tensors produced on the current stream."""
        self._check()
        if self._worker is None:
            self._q = queue.Queue(maxsize=self.depth)
            self._worker = threading.Thread(target=self._spill_loop, daemon=True)
            self._worker.start()
        ev = self.torch.cuda.Event()
        ev.record(self.torch.cuda.current_stream(self.device))
        self.bytes += need
        self.index.add(gidx)
        t0 = time.perf_counter()
        while True:
            try:
                self._q.put((gidx, ev, keep), timeout=1.0)
                self.wait_put_s += time.perf_counter() - t0
                return
            except queue.Full:
                self._check()

    def finish_spill(self):
        if self._worker is not None:
            while True:                      # a worker that died with a full queue must not hang the job
                try:
                    self._q.put(None, timeout=1.0)
                    break
                except queue.Full:
                    self._check()
            self._worker.join()
            self._worker, self._q = None, None
        self.torch.set_num_threads(self._threads0)   # no-op where the setting is per thread
        self._check()

    # -------------------------------------------------------------- pass C
    def _fetch_loop(self, order, compute_stream, slots, stream, outq):
        torch = self.torch
        try:
            torch.cuda.set_device(self.device)
            torch.set_num_threads(self.copy_threads[1])
            for gidx in order:
                tb = time.perf_counter()
                entry = self.store[gidx]
                out, segments = {}, []
                with torch.cuda.stream(stream):
                    for k, hs in entry.items():
                        ts = []
                        for h, dt, shape in hs:
                            t = torch.empty(shape, dtype=dt, device=self.device)
                            t.record_stream(compute_stream)
                            ts.append(t)
                            if t.numel():
                                segments.append((self._bytes(t), h))
                        out[k] = tuple(ts)
                self._pump(segments, False, slots, stream)
                ev = torch.cuda.Event()
                ev.record(stream)
                self.fetch_busy_s += time.perf_counter() - tb
                while True:
                    try:
                        outq.put((gidx, out, ev), timeout=1.0)
                        break
                    except queue.Full:
                        if self._stop:
                            return
                del self.store[gidx]         # the host copy is not needed again
        except BaseException as e:  # noqa: BLE001
            self.err = e

    def start_fetch(self, order):
        order = [g for g in order if g in self.index]
        self._stop = False
        self._pos = {g: i for i, g in enumerate(order)}
        half = len(self.ring) // 2
        cur = self.torch.cuda.current_stream(self.device)
        self._outs, self._workers = [], []
        for w, (slots, stream) in enumerate(((range(0, half), self.cs), (range(half, len(self.ring)), self.cs2))):
            outq = queue.Queue(maxsize=max(1, self.depth // 2))
            th = threading.Thread(target=self._fetch_loop, args=(order[w::2], cur, slots, stream, outq), daemon=True)
            th.start()
            self._outs.append(outq)
            self._workers.append(th)

    def take(self, gidx):
        t0 = time.perf_counter()
        w = self._pos[gidx] % 2
        while True:
            self._check()
            try:
                g, out, ev = self._outs[w].get(timeout=1.0)
                break
            except queue.Empty:
                if not self._workers[w].is_alive():
                    self._check()
                    raise RuntimeError("host tier: prefetcher ended before batch %d" % gidx)
        self.wait_take_s += time.perf_counter() - t0
        if g != gidx:
            raise RuntimeError("host tier: batch %d delivered, %d expected" % (g, gidx))
        self.torch.cuda.current_stream(self.device).wait_event(ev)
        return out

    def close(self):
        self._stop = True
        if self._worker is not None:
            if self._q is not None:
                try:
                    self._q.put_nowait(None)
                except queue.Full:
                    pass                     # daemon thread: left behind after an error, never waited for
            self._worker.join(timeout=30.0)
        for th in self._workers:
            th.join(timeout=30.0)
        self.torch.set_num_threads(self._threads0)
        self._worker, self._q, self._out = None, None, None
        self._workers, self._outs = [], []
        self.store, self.index = {}, set()
        self.bytes = 0


class StreamingSelector:
    def __init__(self, engine, types, cutoff=5.0, n_clusters=10000, n_each=1, seed=0, periodic=True,
                 perceive_bond_orders=False, pool_rows=1 << 21, table_cap=1024, want_molecules=False,
                 only_keys: Optional[Iterable[int]] = None, log: Optional[Callable[[str], None]] = None,
                 kmeans_kwargs: Optional[dict] = None, cache_gb: Optional[float] = None, cache_reserve_gb: float = 24.0,
                 host_cache_gb: Optional[float] = None, host_reserve_gb: float = 24.0,
                 train_threads: int = 1, overlap_seeding: bool = False):
        import torch

        self.eng = engine
        self.types = engine.to_device_types(np.asarray(types).astype(np.uint8))
        self.N = int(self.types.numel())
        self.cutoff = float(cutoff)
        self.K = int(n_clusters)
        self.n_each = int(n_each)
        self.seed = seed
        self.periodic = bool(periodic)
        self.perceive = bool(perceive_bond_orders)
        self.pool_rows = max(int(pool_rows), 3 * self.K)
        self.cap = int(table_cap)
        self.want_molecules = want_molecules
        self.only_keys = None if only_keys is None else {int(k) for k in only_keys}
        self.log = log or (lambda s: None)
        self.kmeans_kwargs = dict(kmeans_kwargs or {})
        self.torch = torch
        self.nt = len(engine.atomname)
        self.comm = {}
        self.cache_gb = cache_gb                  # HBM for spectra kept from pass B to pass C (None: what is free)
        self.cache_reserve_gb = cache_reserve_gb
        self.cache, self.cache_bytes, self.cache_budget, self.cache_hits = {}, 0, 0, 0
        self.host_cache_gb = host_cache_gb        # host memory for spectra beyond the HBM tier (None: what is free, 0: off)
        self.host_reserve_gb = host_reserve_gb
        self.host = None
        self.capture_rows = None
        self.train_threads = int(train_threads)
        self.overlap_seeding = bool(overlap_seeding)
        self._workers = []
        self.eng_profile = False

    # ------------------------------------------------------------------ helpers
    def _sync(self):
        self.torch.cuda.current_stream(self.eng.device).synchronize()

    def _keys_of(self, b: FrameBatch, labels_out=None):
        eng = self.eng
        if b.rowptr is not None:
            keys = eng.bond_keys_csr(b.rowptr, b.bo, self.types)
            if self.want_molecules:
                eng.components_csr(b.rowptr, b.col, out=labels_out)
            return keys
        if b.nbr is not None:
            keys = eng.bond_keys_bo(b.nbr, b.bo, self.types)
            if self.want_molecules:
                eng.components(b.nbr)
            return keys
        eng.set_option("perceive_bond_orders", 1 if self.perceive else 0)
        try:
            out = eng.analyze(b.pos, b.cell, self.types, self.periodic, want_ell=False, want_keys=True,
                              want_labels=self.want_molecules or self.perceive)
        finally:
            eng.set_option("perceive_bond_orders", 0)
        return out["keys"]

    # ------------------------------------------------------------------ pass A
    def pass_a(self, source):
        torch, eng = self.torch, self.eng
        table = eng.key_table(self.cap)
        maxc = torch.zeros((self.cap, eng.MAXT), dtype=torch.int32, device=eng.device)
        stores: List[_Store] = []
        nstep = 0
        rings = 0
        for b in source.batches():
            F = int(b.pos.shape[0])
            keys = self._keys_of(b)
            if b.rowptr is None and b.nbr is None and self.perceive:
                rings += eng.perceive_stats()["rings"]
            codes, order, seg = eng.group_keys(keys, table, b.step0, keep=b.keep)
            del keys
            st = _Store(b.gidx, b.step0, F, order, seg)
            st.kept = F * self.N if b.keep is None else int(seg[self.cap].item())
            if st.kept:
                tg = order[: st.kept]
                counts, _ = eng.compositions(b.pos, b.cell, self.types, self.cutoff, targets=tg, periodic=self.periodic)
                eng.type_maxcount(codes.view(-1), counts, maxc, targets=tg)
                del counts
            if st.kept != F * self.N:
                st.order = order[: st.kept].clone()
            del codes
            stores.append(st)
            nstep += F
        eng.check()
        self.rings = rings
        segs = torch.stack([s.seg_d for s in stores]).cpu().numpy() if stores else np.zeros((0, self.cap + 1), np.int64)
        for s, sg in zip(stores, segs):
            s.seg = sg
            s.seg_d = None
        tk = table[0].cpu().numpy().view(np.uint64)
        tc = table[1].cpu().numpy()
        tf = table[2].cpu().numpy().view(np.uint64)
        mc = maxc.cpu().numpy()
        slots = np.nonzero(tk != KEY_EMPTY)[0]
        local = {int(tk[s]): (int(tc[s]), int(tf[s]), mc[s, : self.nt].copy()) for s in slots}
        self.slot_of = {int(tk[s]): int(s) for s in slots}
        # per-batch row counts of every key: [(gidx, {key: n})]
        per_batch = [(st.gidx, {int(tk[s]): int(st.seg[s + 1] - st.seg[s]) for s in slots if st.seg[s + 1] > st.seg[s]})
                     for st in stores]
        d = _dist()
        if d is not None:
            gathered = [None] * d.get_world_size()
            d.all_gather_object(gathered, (local, per_batch, nstep))
            merged: Dict[int, list] = {}
            per_batch, nstep = [], 0
            for loc, pb, ns in gathered:
                for k, (cnt, first, m) in loc.items():
                    if k in merged:
                        merged[k][0] += cnt
                        merged[k][1] = min(merged[k][1], first)
                        merged[k][2] = np.maximum(merged[k][2], m)
                    else:
                        merged[k] = [cnt, first, m.copy()]
                per_batch.extend(pb)
                nstep += ns
            local = {k: tuple(v) for k, v in merged.items()}
        self.stores = stores
        self.nstep = nstep
        self.keys = sorted(local, key=lambda k: local[k][1])        # first-seen order (datasetbuilder.py:205-206)
        self.counts = {k: local[k][0] for k in self.keys}
        self.maxcount = {k: local[k][2].astype(np.int32) for k in self.keys}
        per_batch.sort(key=lambda x: x[0])
        self.gbatches = [g for g, _ in per_batch]
        self.gpos = {g: i for i, g in enumerate(self.gbatches)}
        # rows of key k before global batch i: row0[k][i]
        self.row0 = {}
        for k in self.keys:
            c = np.array([pb.get(k, 0) for _, pb in per_batch], dtype=np.int64)
            self.row0[k] = np.concatenate([[0], np.cumsum(c)])
        self.big = [k for k in self.keys if self.counts[k] > self.K and (self.only_keys is None or k in self.only_keys)]
        allmax = max((int(v[2].sum()) for v in local.values()), default=1)
        self.memcap = max(8, (allmax + 7) // 8 * 8)
        return self

    # ------------------------------------------------------------------ rows of one batch, type by type
    def _ranges(self, st: _Store):
        ranges = []
        for k in self.big:
            s = self.slot_of.get(k)
            if s is None:
                continue
            a, e = int(st.seg[s]), int(st.seg[s + 1])
            if e > a:
                ranges.append((k, a, e))
        return ranges

    def _batch_rows(self, b: FrameBatch, st: _Store, fill_cache=False):
        """Yields (key, X [rows, D], a, e) for every clustered type present in the batch: rows a .. e of the
        batch's order array.  Spectra kept by pass B (`fill_cache`, while the HBM budget lasts) are reused by
        pass C instead of a second sphere scan + eigen-solve: `mdb_feature_rows_compact` rebuilds the rows
        bit for bit."""
        eng, torch = self.eng, self.torch
        if not st.kept or not self.big:
            return
        ranges = self._ranges(st)
        if not ranges:
            return
        cached = self.cache.get(st.gidx)
        if cached is None and not fill_cache and self.host is not None and st.gidx in self.host:
            cached = self.host.take(st.gidx)
        if cached is not None and not fill_cache:
            for k, a, e in ranges:
                eig, off, cnt = cached[k]
                yield k, eng.feature_rows_compact(eig, off, cnt, self.maxcount[k]), a, e
            if self.cache.pop(st.gidx, None) is not None:   # pass C visits a batch once: its HBM goes back right away
                self.cache_hits += 1
            return
        lo, hi = min(a for _, a, _ in ranges), max(e for _, _, e in ranges)
        tg = st.order[lo:hi]
        while True:
            counts, _, members = eng.compositions(b.pos, b.cell, self.types, self.cutoff, targets=tg,
                                                  periodic=self.periodic, members_cap=self.memcap)
            try:
                eng.check()
                break
            except RuntimeError as e:                      # (2) = a membership list overflowed its capacity
                if "(2)" not in str(e) or self.memcap >= 512:
                    raise
                self.memcap = min(512, self.memcap * 2)
        keep, spill = None, False
        # sphere sizes summed per type: one read-back per batch
        csum = torch.cat([torch.zeros(1, dtype=torch.int64, device=eng.device), counts.sum(dim=1).cumsum(0)])
        ends = csum[torch.as_tensor([x - lo for _, a, e in ranges for x in (a, e)] + [hi - lo], device=eng.device)].cpu().numpy()
        first = {k: int(ends[2 * i]) for i, (k, _, _) in enumerate(ranges)}
        totals = {k: int(ends[2 * i + 1] - ends[2 * i]) for i, (k, _, _) in enumerate(ranges)}
        del csum
        if fill_cache and (self.cache_budget > 0 or self.host is not None):
            need = sum(8 * totals[k] + (e - a) * (4 + 4 * self.nt) + 1024 for k, a, e in ranges)
            to_host = self.host is not None and self.host.room(need)
            to_hbm = need <= self.cache_budget
            if to_hbm and to_host:
                # both tiers have room: spread the HBM share evenly over the batches instead of filling HBM first, so
                # that pass C meets host batches between HBM batches and the prefetchers get time to run ahead
                if self._hbm_share is None:
                    self._hbm_share = min(1.0, self.cache_budget / float(need * max(1, len(self.stores))))
                self._hbm_acc += self._hbm_share
                to_hbm = self._hbm_acc >= 1.0
                if to_hbm:
                    self._hbm_acc -= 1.0
            if to_hbm:
                self.cache_budget -= need
                self.cache_bytes += need
                keep = {}
            elif to_host:
                keep, spill = {}, True
        # ONE eigen-solve call for every type of the batch (the spectrum of a sphere does not depend on the type it is
        # filed under): one size histogram and one read-back instead of one per type, and launches that fill the GPU
        # also for the rare types.  Each type's slice of the compact spectra is copied out (the long-lived tensors then
        # have the same sizes as with one call per type, which the allocator handles without fragmenting -- views of
        # the whole buffer did not survive the full-size job) and its rows are merged from the slice.
        _, eig_all, off_all = eng.descriptors_compact(b.pos, b.cell, self.types, counts, self.maxcount[ranges[0][0]], tg, members,
                                                      int(ends[-1]), periodic=self.periodic, want_X=False)
        del members
        parts = {}
        for k, a, e in ranges:
            parts[k] = (eig_all[first[k]:first[k] + totals[k]].clone(), off_all[a - lo:e - lo + 1] - off_all[a - lo],
                        counts[a - lo:e - lo].clone())
        del eig_all, off_all, counts
        for k, a, e in ranges:
            eig, off, cn = parts.pop(k)
            X = eng.feature_rows_compact(eig, off, cn, self.maxcount[k])
            if keep is not None:
                keep[k] = (eig, off, cn)
            del eig, off, cn
            yield k, X, a, e
        if spill:
            self.host.put(st.gidx, keep, need)
        elif keep is not None:
            self.cache[st.gidx] = keep

    # ------------------------------------------------------------------ pass B
    def _pool_indices(self):
        self.pool_idx = {}
        for k in self.big:
            n = self.counts[k]
            if n <= self.pool_rows:
                self.pool_idx[k] = None                   # the pool is every row
            else:
                rng = np.random.default_rng([int(self.seed) & 0xFFFFFFFF, k & 0xFFFFFFFF, k >> 32])
                self.pool_idx[k] = np.sort(rng.choice(n, size=self.pool_rows, replace=False, shuffle=False))

    def pass_b(self, source):
        torch, eng = self.torch, self.eng
        self._pool_indices()
        self.D = {k: int(self.maxcount[k].sum()) for k in self.big}
        big_val = np.finfo(np.float64).max
        self.mn = {k: torch.full((self.D[k],), big_val, dtype=torch.float64, device=eng.device) for k in self.big}
        self.mx = {k: torch.full((self.D[k],), -big_val, dtype=torch.float64, device=eng.device) for k in self.big}
        self.pool = {k: torch.zeros((min(self.counts[k], self.pool_rows), self.D[k]), dtype=torch.float64,
                                    device=eng.device) for k in self.big}
        stores = {s.gidx: s for s in self.stores}
        # spectra cache for pass C: what is free now minus the labels pass C stores and its working set
        self.cache, self.cache_bytes, self.cache_hits = {}, 0, 0
        self._hbm_share, self._hbm_acc = None, 0.0
        if self.cache_gb is None:
            free, _ = torch.cuda.mem_get_info(eng.device)
            free += torch.cuda.memory_reserved(eng.device) - torch.cuda.memory_allocated(eng.device)
            rows_local = sum(int(st.kept) for st in self.stores)
            self.cache_budget = int(free - 4 * rows_local - self.cache_reserve_gb * (1 << 30))
        else:
            self.cache_budget = int(self.cache_gb * (1 << 30))
        # second tier: host memory (this rank's share of what is available minus a reserve)
        if self.host is not None:
            self.host.close()
        self.host = None
        if self.host_cache_gb is None:
            share = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))
            hb = (_host_memory_available() - int(self.host_reserve_gb * (1 << 30))) // share
        else:
            hb = int(self.host_cache_gb * (1 << 30))
        if hb > ((64 << 20) if self.host_cache_gb is None else 0):
            self.host = _HostTier(torch, eng.device, hb)
            if self.cache_gb is None:
                self.cache_budget -= 8 << 30      # batches queued for the spill / prefetched in pass C live in HBM too
        for b in source.batches():
            st = stores[b.gidx]
            gi = self.gpos[b.gidx]
            for k, X, a, e in self._batch_rows(b, st, fill_cache=True):
                if self.capture_rows is not None:      # tests: the unscaled rows of every type, in row order
                    self.capture_rows.setdefault(k, []).append((b.gidx, X.clone()))
                eng.minmax_accumulate(X, self.mn[k], self.mx[k])
                g0 = int(self.row0[k][gi])
                g1 = g0 + (e - a)
                pidx = self.pool_idx[k]
                if pidx is None:
                    src = torch.arange(0, e - a, dtype=torch.int64, device=eng.device)
                    dst = src + g0
                else:
                    l, h = np.searchsorted(pidx, [g0, g1])
                    if h <= l:
                        continue
                    src = torch.as_tensor(pidx[l:h] - g0, dtype=torch.int64, device=eng.device)
                    dst = torch.arange(l, h, dtype=torch.int64, device=eng.device)
                eng.gather_rows(X, src, dst, self.pool[k])
                del X
        eng.check()
        if self.host is not None:
            self.host.finish_spill()
        d = _dist()
        if d is not None:
            self._sync()
            t0 = time.perf_counter()
            nbytes = 0
            for k in self.big:
                d.all_reduce(self.mn[k], op=d.ReduceOp.MIN)
                d.all_reduce(self.mx[k], op=d.ReduceOp.MAX)
                d.all_reduce(self.pool[k])                # every row is written by exactly one rank, zeros elsewhere
                nbytes += self.pool[k].numel() * 8 + 16 * self.D[k]
            self._sync()
            self.comm["pool_and_bounds"] = {"ms": (time.perf_counter() - t0) * 1e3, "bytes": int(nbytes)}
        self.mn_h = {k: self.mn[k].cpu().numpy() for k in self.big}
        self.mx_h = {k: self.mx[k].cpu().numpy() for k in self.big}
        return self

    # ------------------------------------------------------------------ training
    def _fit_one(self, eng, k):
        pool = self.pool[k]
        eng.minmax_transform(pool, self.mn_h[k], self.mx_h[k])
        n = int(pool.shape[0])
        km = MiniBatchKMeansB200(eng, n_clusters=self.K, init_size=min(3 * self.K, n), n_init=3, seed=self.seed,
                                 **self.kmeans_kwargs)
        km.fit(pool, compute_labels=False)
        stats = {"steps": km.n_steps_, "pool_rows": n, "rows": self.counts[k], "D": self.D[k]}
        return km.cluster_centers_, stats, km

    def train(self):
        """MiniBatchKMeans per clustered type on its scaled pool.  A fit is a seeding (3 x k-means++, one long
        handshake-bound kernel) followed by a loop of small mini-batch steps with a host decision after each;
        the seeding of the next type runs on a second context / stream while the loop of the current one runs
        (`overlap_seeding`).  With several ranks the types are dealt over the ranks and the centres broadcast by
        their owner.  Every fit carries its own RandomState, so neither changes any result.  The "allreduce"
        exchange keeps every rank in every fit instead (mini-batch slices + one packed all-reduce per step).
        Measured on B200 (15 types, K = 10,000): the overlap is a loss -- the persistent seeding kernel spins on its
        flags and the step kernels of the other stream starve each other (train 8.2 -> 17.4 s) -- so it is off by
        default.  (Also tried and dropped: several fits on host threads -- the co-operative seeding kernels serialise and the
        interpreter lock makes the step loops slower, 8.2 -> 9.5 s for 15 types.)"""
        torch, eng = self.torch, self.eng
        self.centers = {}
        self.fit_stats = {}
        self.fit_comm = {"allreduce_ms": 0.0, "steps": 0, "bytes_per_step": 0}
        d = _dist()
        world = d.get_world_size() if d is not None else 1
        rank = d.get_rank() if d is not None else 0
        todo = sorted(self.big, key=lambda k: (-int(self.pool[k].shape[0]) * self.D[k], k))
        if self.kmeans_kwargs.get("exchange") == "allreduce" and world > 1:
            for k in self.big:
                self.centers[k], self.fit_stats[k], km = self._fit_one(eng, k)
                fc = self.fit_comm
                fc["allreduce_ms"] += km.comm_ms_ or 0.0
                fc["steps"] += km.n_steps_
                fc["bytes_per_step"] = max(fc["bytes_per_step"], getattr(km, "comm_bytes_per_step_", 0))
            self.pool = None
            return self
        owner = {k: i % world for i, k in enumerate(todo)}
        mine = [k for k in todo if owner[k] == rank]
        self._sync()
        if not self.overlap_seeding or len(mine) < 2:
            for k in mine:
                self.centers[k], self.fit_stats[k], _ = self._fit_one(eng, k)
        else:
            # seeding (scaling of the pool, validation draw, 3 x k-means++: one long co-operative kernel that is
            # bound by its own handshakes) of type t + 1 on a second context / stream while the mini-batch loop of
            # type t runs on the main one: the seeds arrive through a queue, each fit keeps its own RandomState
            import queue
            import threading

            from .engine import Engine

            if not self._workers:
                self._workers.append((Engine(eng.device.index, eng.atomname), torch.cuda.Stream(device=eng.device)))
            w, stream = self._workers[0]
            w.profile(getattr(eng, "_profiling", False))
            q: "queue.Queue" = queue.Queue(maxsize=2)

            def seeder():
                try:
                    torch.cuda.set_device(eng.device)
                    with torch.cuda.stream(stream):
                        for k in mine:
                            pool = self.pool[k]
                            w.minmax_transform(pool, self.mn_h[k], self.mx_h[k])
                            n = int(pool.shape[0])
                            km = MiniBatchKMeansB200(w, n_clusters=self.K, init_size=min(3 * self.K, n), n_init=3,
                                                     seed=self.seed, **self.kmeans_kwargs)
                            state = km.seed_centers(pool)
                            stream.synchronize()
                            q.put((k, state))
                except Exception as e:          # surfaced on the main thread
                    q.put((None, e))

            th = threading.Thread(target=seeder)
            th.start()
            try:
                for _ in mine:
                    k, state = q.get()
                    if k is None:
                        raise state
                    pool = self.pool[k]
                    n = int(pool.shape[0])
                    km = MiniBatchKMeansB200(eng, n_clusters=self.K, init_size=min(3 * self.K, n), n_init=3, seed=self.seed,
                                             **self.kmeans_kwargs)
                    km.run_loop(pool, state)
                    self.centers[k] = km.cluster_centers_
                    self.fit_stats[k] = {"steps": km.n_steps_, "pool_rows": n, "rows": self.counts[k], "D": self.D[k]}
                    self.pool[k] = None
            finally:
                th.join()
        if world > 1:
            self._sync()
            t0 = time.perf_counter()
            for k in todo:
                if owner[k] != rank:
                    self.centers[k] = torch.empty((self.K, self.D[k]), dtype=torch.float64, device=eng.device)
                d.broadcast(self.centers[k], src=owner[k])
            gathered = [None] * world
            d.all_gather_object(gathered, self.fit_stats)
            for part in gathered:
                self.fit_stats.update(part)
            self._sync()
            self.comm["centres_broadcast_incl_wait_for_slowest_rank"] = {"ms": (time.perf_counter() - t0) * 1e3,
                                              "bytes": int(sum(self.K * self.D[k] * 8 for k in todo))}
        self.pool = None
        return self

    def worker_engines(self):
        return [w for w, _ in self._workers]

    # ------------------------------------------------------------------ pass C
    def pass_c(self, source):
        torch, eng = self.torch, self.eng
        nb = len(self.gbatches)
        self.labels = {k: [] for k in self.big}           # per local batch: (gidx, labels tensor)
        self.hist = {k: torch.zeros((nb, self.K), dtype=torch.int32, device=eng.device) for k in self.big}
        stores = {s.gidx: s for s in self.stores}
        rechecked = 0
        # batches whose spectra were kept need no input tensors: sources that stage data (host -> device copies,
        # text parsing) may skip them
        host = self.host
        if host is not None and len(host):
            host.start_fetch([g for g in self.gbatches if g in stores])
        try:
            it = source.batches(needed=lambda g: g not in self.cache and (host is None or g not in host))
        except TypeError:
            it = source.batches()
        for b in it:
            st = stores[b.gidx]
            gi = self.gpos[b.gidx]
            for k, X, a, e in self._batch_rows(b, st):
                eng.minmax_transform(X, self.mn_h[k], self.mx_h[k])
                if self.D[k] <= 112 and (e - a) >= 4096:
                    lab, nre = eng.kmeans_assign_tc(X, self.centers[k])
                    rechecked += nre
                else:
                    lab, _, _ = eng.kmeans_assign(X, self.centers[k])
                eng.label_hist(lab, self.hist[k][gi])
                self.labels[k].append((b.gidx, a, e, lab))
                del X
        eng.check()
        self.cached_batches = self.cache_hits
        self.cache = {}
        self.host_batches, self.host_bytes = (len(host), host.bytes) if host is not None else (0, 0)
        if host is not None:
            self.host_waits = {"pass_b_queue_full_s": round(host.wait_put_s, 3), "pass_c_waiting_for_batch_s": round(host.wait_take_s, 3),
                               "spill_worker_busy_s": round(host.spill_busy_s, 3), "fetch_worker_busy_s": round(host.fetch_busy_s, 3)}
            host.close()
        self.rechecked = rechecked
        d = _dist()
        if d is not None:
            self._sync()
            t0 = time.perf_counter()
            for k in self.big:
                d.all_reduce(self.hist[k])                # a batch belongs to one rank
            self._sync()
            self.comm["cluster_counts"] = {"ms": (time.perf_counter() - t0) * 1e3,
                                           "bytes": int(sum(self.hist[k].numel() * 4 for k in self.big))}
        return self

    # ------------------------------------------------------------------ selection
    def select(self) -> SelectionResult:
        torch, eng = self.torch, self.eng
        d = _dist()
        stores = {s.gidx: s for s in self.stores}
        chosen: Dict[int, list] = {}
        N = self.N
        for k in self.keys:
            if self.only_keys is not None and k not in self.only_keys:
                continue
            if k in self.big:
                H = self.hist[k].cpu().numpy().astype(np.int64)          # [global batches, K]
                totals = H.sum(axis=0)
                nz = np.nonzero(totals)[0]
                rng = np.random.default_rng(self.seed)
                draws = rng.integers(0, totals[nz][:, None], size=(len(nz), self.n_each))   # as choose_per_cluster
                cum = np.cumsum(H[:, nz], axis=0)                            # [nb, nnz]
                mine = []
                if self.labels[k]:
                    lab_all = torch.cat([t[3] for t in self.labels[k]])
                    base, off = {}, 0
                    for g, a, e, t in self.labels[k]:
                        base[g] = (off, off + (e - a), a)
                        off += e - a
                    q, tag = [], []
                    for j in range(self.n_each):
                        p = draws[:, j]
                        gi = (cum <= p[None, :]).sum(axis=0)                  # global batch of the p-th member
                        prev = np.where(gi > 0, cum[np.maximum(gi - 1, 0), np.arange(len(nz))], 0)
                        for c in range(len(nz)):
                            g = self.gbatches[gi[c]]
                            if g in base:
                                r0, r1, _ = base[g]
                                q.append((r0, r1, int(nz[c]), int(p[c] - prev[c])))
                                tag.append((int(nz[c]), j, g))
                    if q:
                        rows = eng.pick_rows(lab_all, np.asarray(q, dtype=np.int64)).cpu().numpy()
                        if (rows < 0).any():
                            raise RuntimeError("selection: a cluster holds fewer members than counted (internal error)")
                        # row -> atom-frame through the batch's order array
                        by_g: Dict[int, list] = {}
                        for r, (kk, j, g) in zip(rows, tag):
                            by_g.setdefault(g, []).append((int(r), kk, j))
                        for g, lst in by_g.items():
                            r0, _, a = base[g]
                            st = stores[g]
                            idx = torch.as_tensor([a + (r - r0) for r, _, _ in lst], dtype=torch.int64, device=eng.device)
                            af = st.order[idx].cpu().numpy().astype(np.int64)
                            for v, (_, kk, j) in zip(af, lst):
                                mine.append((st.step0 + int(v // N), int(v % N) + 1, (kk, j)))
                    del lab_all
            else:
                mine = []
                s = self.slot_of.get(k)
                if s is not None:
                    for st in self.stores:
                        a, e = int(st.seg[s]), int(st.seg[s + 1])
                        if e > a:
                            af = st.order[a:e].cpu().numpy().astype(np.int64)
                            mine.extend((st.step0 + int(v // N), int(v % N) + 1, (st.step0 + int(v // N), int(v % N) + 1))
                                        for v in af)
            if d is not None:
                gathered = [None] * d.get_world_size()
                d.all_gather_object(gathered, mine)
                mine = [c for part in gathered for c in part]
            mine.sort(key=lambda c: c[2])
            chosen[k] = [(c[0], c[1]) for c in mine]
        return SelectionResult(self.keys, self.counts, chosen, {k: self.maxcount[k] for k in self.big}, self.nstep)

    # ------------------------------------------------------------------ whole job
    def run(self, source) -> SelectionResult:
        t = [time.perf_counter()]

        def lap():
            self._sync()
            t.append(time.perf_counter())
            return t[-1] - t[-2]

        tm = {}
        self.pass_a(source)
        tm["pass_a"] = lap()
        self.log(f"pass A: {self.nstep} frames, {len(self.keys)} bond types, {len(self.big)} to cluster")
        if self.big:
            self.pass_b(source)
            tm["pass_b"] = lap()
            self.train()
            tm["train"] = lap()
            self.pass_c(source)
            tm["pass_c"] = lap()
        res = self.select()
        tm["select"] = lap()
        res.timings = tm
        res.stats = {"fit": getattr(self, "fit_stats", {}), "rechecked_in_float64": getattr(self, "rechecked", 0),
                     "memcap": self.memcap, "rings": getattr(self, "rings", 0), "comm": self.comm,
                     "spectra_cache": {"batches": getattr(self, "cached_batches", 0), "of": len(self.stores),
                                       "gb": self.cache_bytes / 1e9, "host_batches": getattr(self, "host_batches", 0),
                                       "host_gb": getattr(self, "host_bytes", 0) / 1e9, "host_waits": getattr(self, "host_waits", None)},
                     "hbm_peak_allocated_gb": self.torch.cuda.max_memory_allocated(self.eng.device) / 1e9,
                     "allocator": {k: int(v) for k, v in self.torch.cuda.memory_stats(self.eng.device).items()
                                   if k in ("num_alloc_retries", "num_ooms", "segment.all.allocated", "segment.all.freed",
                                            "num_device_alloc", "num_device_free")},
                     "fit_comm": getattr(self, "fit_comm", None)}
        return res


class ResidentSource:
    """Frames that already live in HBM (bench.py, tests): positions [F, N, 3], one cell or [F, 9], optional
    CSR bond table (rowptr [F, N+1] / [N+1] shared, col, bo [F, colcap]) or ELL table (nbr, bo [F, N, width]).
    Batches of ``frames_per_batch`` frames are dealt round-robin to the ranks like DatasetBuilder deals file
    batches; with ``local=True`` every rank owns all of its frames (weak scaling) and they interleave."""

    def __init__(self, pos, cell, frames_per_batch, rowptr=None, col=None, bo=None, nbr=None, keep=None,
                 rank=0, world=1, local=False, step_stride=1):
        self.pos, self.rowptr, self.col, self.bo, self.nbr, self.keep = pos, rowptr, col, bo, nbr, keep
        F = int(pos.shape[0])
        c = np.asarray(cell, dtype=np.float64)
        if c.size == 3:
            c = np.diag(c.reshape(3))
        self.cell = np.ascontiguousarray(np.broadcast_to(c.reshape(-1, 9), (F, 9)))
        self.B = int(frames_per_batch)
        self.rank, self.world, self.local = rank, world, local

    def batches(self, needed=None):
        F = int(self.pos.shape[0])
        nb = -(-F // self.B)
        for i in range(nb):
            if self.local:
                g = i * self.world + self.rank
                step0 = g * self.B
            else:
                if i % self.world != self.rank:
                    continue
                g = i
                step0 = i * self.B
            f0, f1 = i * self.B, min(F, (i + 1) * self.B)
            kw = {}
            if self.rowptr is not None:
                rp = self.rowptr
                kw = dict(rowptr=rp[f0:f1], col=self.col[f0:f1], bo=self.bo[f0:f1])
            elif self.nbr is not None:
                kw = dict(nbr=self.nbr[f0:f1], bo=self.bo[f0:f1])
            yield FrameBatch(g, step0, self.pos[f0:f1], self.cell[f0:f1],
                             keep=None if self.keep is None else self.keep[f0:f1], **kw)
