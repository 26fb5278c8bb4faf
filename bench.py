#!/usr/bin/env python
"""Benchmark of the B200-native per-frame analysis path (contract: see the task brief).

    python bench.py --gpus N --steps K --warmup W            # this framework
    python bench.py --impl reference --gpus N --steps K ...  # CPU baseline arm

Workload (BASELINE.json configs[2], "C3"): synthetic 1M-atom periodic C/H/O box at 0.05 atoms/A^3,
1,000 frames per rank, bond table (CSR with bond orders, the content of a `fix reaxff/bonds` file) +
per-atom Coulomb-matrix descriptor + MiniBatchKMeans(K = 10,000) per bond type + per-cluster selection:
the whole of `DatasetBuilder._readtimestepsbond` + `_writecoulumbmatrix` for every type
(reference datasetbuilder.py:185-281, 351-384), run as the streaming job of
`mddatasetbuilder_b200/streaming.py` (pass A keys / group-by / compositions / molecule labels, pass B
descriptors -> scaler bounds + training pool, training, pass C descriptors -> scale -> assignment,
selection).  The job's frames are split into K equal steps (a step = frames/K frames taken through all
three passes; training and selection happen once per job and are inside the timed region), so
`value` = atom-frames of the job / wall time of the WHOLE job.  Warm-up = a smaller job of W steps.
With N > 1 (torchrun) every rank owns 1,000 frames (configs[3] shape, weak scaling): the trajectory is
the union, batches interleave over the ranks, the collectives of the job are inside the timed region.

`value` has the inputs resident in HBM; `e2e` runs the same job from page-locked host arrays (every
batch of every pass is copied host -> device inside the timed region, the selection comes back to the
host).  `stages.c2_bonds` keeps the dump-only bond path of configs[1] (round 1's headline).
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

C3_TEXT = ("configs[2]: synthetic 1M-atom periodic C/H/O box, 1,000 frames, bonds file + per-atom descriptor + "
           "MiniBatchKMeans, 1 B200")
C4_TEXT = ("configs[3] shape: synthetic 1M-atom trajectory, 1,000 frames per rank sharded across the B200s "
           "(weak scaling of configs[2]), NCCL collectives of the clustering inside the timed region")
METRIC = "atom-frames/sec (bond+descriptor+kmeans) at 1/2/4/8 B200; HBM GB/s vs peak"
UNIT = "atom-frames/s"
NATOMS, FRAMES, DENSITY, KCLUST = 1_000_000, 1000, 0.05, 10_000


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "250", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.3)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ts, line in self.lines:
            if ts < t0 - 0.05 or ts > t1 + 0.3:
                continue
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples inside the timed region"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------------------
# CPU baseline: the oracle port of the reference's per-frame stages on the host cores
# --------------------------------------------------------------------------------------
def cpu_baseline(natoms=NATOMS, density=DENSITY, seed=20261019 + 3, target_seconds=15.0, cores=None, sysm=None):
    """The reference's CPU path for this workload, stage by stage, on ONE 1M-atom frame with every
    host thread busy, each stage on a bounded sample (kind "port": oracle/ restates ASE's minimum-image
    sphere cut and the Coulomb matrix in C, numpy.linalg.eigh and scikit-learn are the real ones):
      keys        oracle bond_keys_bondfile on the whole frame (detect.py:105-119), one thread;
      descriptor  `_writestepmatrix` + `_calcoulumbmatrix` (datasetbuilder.py:283-349): O(N) sphere cut
                  against all 1M atoms + Coulomb matrix + eigh, S sampled targets per thread;
      assignment  sklearn `_labels_inertia` of those rows against K = 10,000 centres.
    value = cores / (seconds per atom-frame summed over the stages)."""
    from concurrent.futures import ThreadPoolExecutor

    from mddatasetbuilder_b200 import synth
    from oracle import oracle as O

    cores = cores or len(os.sched_getaffinity(0))
    s = sysm or synth.make_system(natoms, density, seed=seed)
    Z = np.array([O.ATOMIC_NUMBERS[s.atomname[t]] for t in s.types], dtype=np.int32)
    diag = np.array([O.ATOMIC_NUMBERS[s.atomname[t]] ** 2.4 / 2 for t in s.types])
    t0 = time.perf_counter()
    O.bond_keys_bondfile(s.types, s.rowptr, s.bo)
    t_keys = (time.perf_counter() - t0) / natoms               # seconds per atom-frame, one thread
    # calibrate the descriptor sample: a handful of targets on one thread
    rng = np.random.default_rng(seed)
    probe = rng.integers(0, natoms, 4).astype(np.int32)
    t0 = time.perf_counter()
    O.descriptor_gather(Z, diag, s.pos0, s.box, True, 5.0, probe)
    per_target = (time.perf_counter() - t0) / len(probe)
    S = max(2, int(target_seconds / max(per_target, 1e-6)))

    def work(k):
        tg = np.random.default_rng(seed + 1 + k).integers(0, natoms, S).astype(np.int32)
        counts, mem, mats = O.descriptor_gather(Z, diag, s.pos0, s.box, True, 5.0, tg)
        eigs = O.coulomb_eigs(mats)
        return counts, eigs, mem

    t0 = time.perf_counter()
    with ThreadPoolExecutor(cores) as ex:
        outs = list(ex.map(work, range(cores)))
    wall_desc = time.perf_counter() - t0
    t_desc = wall_desc / S                                      # seconds per atom-frame per thread
    # assignment of the sampled rows against K centres with the real scikit-learn (its own thread pool)
    eigs = [e for _, es, _ in outs for e in es]
    ec = np.array([[int((s.types[m] == el).sum()) for el in range(len(s.atomname))] for _, _, ms in outs for m in ms])
    pads = np.array([O.ATOMIC_NUMBERS[a] ** 2.4 / 2 for a in s.atomname])
    X, _ = O.feature_rows(eigs, ec, pads)
    Xs = O.minmax_scale(X)[0]
    reps = max(1, 20000 // len(Xs))
    Xa = np.tile(Xs, (reps, 1))
    C = Xa[np.random.default_rng(seed).integers(0, len(Xa), KCLUST)] + 1e-3
    from sklearn.cluster._kmeans import _labels_inertia

    t0 = time.perf_counter()
    _labels_inertia(np.ascontiguousarray(Xa), np.ones(len(Xa)), np.ascontiguousarray(C), n_threads=cores)
    t_assign = (time.perf_counter() - t0) / len(Xa) * cores     # core-seconds per row
    per_atom_frame = t_keys + t_desc + t_assign
    value = cores / per_atom_frame
    return {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"one {natoms}-atom frame at {density} atoms/A^3: bond-type keys of the whole frame (1 thread, "
                      f"{t_keys * 1e9:.0f} ns/atom); sphere cut against all atoms + Coulomb matrix (oracle/oracle.c) + "
                      f"numpy eigh for {S} sampled targets on each of {cores} threads ({t_desc * 1e3:.2f} ms/target/thread, "
                      f"{wall_desc:.1f} s wall); sklearn _labels_inertia of {len(Xa)} rows vs K={KCLUST} "
                      f"({t_assign * 1e6:.1f} core-us/row). Training and selection excluded (negligible next to the "
                      f"O(N) sphere cut per target)",
            "seconds": wall_desc, "targets_per_thread": S,
            "seconds_per_atom_frame_per_core": {"keys": t_keys, "descriptor": t_desc, "assignment": t_assign}}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from mddatasetbuilder_b200 import synth

    sysm = synth.make_system(NATOMS, DENSITY, seed=synth.BASE_SEED + 3)
    per_step = max(2.0, min(12.0, 150.0 / max(args.steps + args.warmup, 1)))
    for _ in range(args.warmup):
        cpu_baseline(target_seconds=per_step / 4, sysm=sysm)
    vals, secs, base = [], [], None
    for _ in range(args.steps):
        t0 = time.perf_counter()
        base = cpu_baseline(target_seconds=per_step, sysm=sysm)
        secs.append(time.perf_counter() - t0)
        vals.append(base["value"])
    value = float(np.mean(vals))
    base["value"] = value
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(secs)), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": {"workload": C3_TEXT, "atoms_per_frame": NATOMS, "frames_per_step": "bounded sample",
                       "note": "reference CPU path = oracle port of its stages (ase / openbabel are not installable "
                               "offline) + the real numpy eigh and scikit-learn; each step is a bounded sample of one "
                               "1M-atom frame on all host threads"},
            "cpu_baseline": base,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


# --------------------------------------------------------------------------------------
class HostSource:
    """The same frames as page-locked host arrays: every batch of every pass is copied to the device inside
    the timed region.  `slots` distinct frames are kept on the host and reused cyclically (frame f reads
    slot f % slots), so the page-locked footprint stays bounded; the bytes copied are the full job's."""

    def __init__(self, torch, dev, pos, rowptr, col, bo, cell, frames, frames_per_batch, rank, world, slots):
        self.torch, self.dev = torch, dev
        self.slots = min(slots, int(pos.shape[0]))
        self.h_pos = torch.empty((self.slots,) + tuple(pos.shape[1:]), dtype=pos.dtype, pin_memory=True)
        self.h_rowptr = torch.empty((self.slots,) + tuple(rowptr.shape[1:]), dtype=rowptr.dtype, pin_memory=True)
        self.h_col = torch.empty((self.slots,) + tuple(col.shape[1:]), dtype=col.dtype, pin_memory=True)
        self.h_bo = torch.empty((self.slots,) + tuple(bo.shape[1:]), dtype=bo.dtype, pin_memory=True)
        self.h_pos.copy_(pos[: self.slots])
        self.h_rowptr.copy_(rowptr[: self.slots])
        self.h_col.copy_(col[: self.slots])
        self.h_bo.copy_(bo[: self.slots])
        self.cell = np.ascontiguousarray(np.broadcast_to(np.asarray(cell, dtype=np.float64).reshape(-1, 9), (frames, 9)))
        self.F, self.B, self.rank, self.world = frames, frames_per_batch, rank, world
        self.h2d_bytes = 0
        B = frames_per_batch
        self.d = [tuple(torch.empty((B,) + tuple(t.shape[1:]), dtype=t.dtype, device=dev)
                        for t in (self.h_pos, self.h_rowptr, self.h_col, self.h_bo)) for _ in range(2)]

    def batches(self, needed=None):
        from mddatasetbuilder_b200.streaming import FrameBatch

        nb = -(-self.F // self.B)
        for i in range(nb):
            f0, f1 = i * self.B, min(self.F, (i + 1) * self.B)
            nf = f1 - f0
            buf = self.d[i & 1]
            g = i * self.world + self.rank
            if needed is not None and not needed(g):              # pass C reuses the spectra it kept: no copy
                yield FrameBatch(g, g * self.B, buf[0][:nf], self.cell[f0:f1], rowptr=buf[1][:nf], col=buf[2][:nf], bo=buf[3][:nf])
                continue
            for f in range(nf):                                   # slot-wise copies (cyclic reuse of the host frames)
                sl = (f0 + f) % self.slots
                for dst, src in zip(buf, (self.h_pos, self.h_rowptr, self.h_col, self.h_bo)):
                    dst[f].copy_(src[sl], non_blocking=True)
                    self.h2d_bytes += src[sl].numel() * src.element_size()
            yield FrameBatch(g, g * self.B, buf[0][:nf], self.cell[f0:f1], rowptr=buf[1][:nf], col=buf[2][:nf], bo=buf[3][:nf])


def measure_c2_bonds(torch, dev, local, max_over_ranks, world, peaks):
    """Round 1's headline, kept as a stage: configs[1], 100k atoms x 1,000 frames, dump-only bond path
    (cell list -> bonds -> cleanup -> keys -> molecule labels), device resident, 3 warm-up + 5 timed passes."""
    from mddatasetbuilder_b200 import synth
    from mddatasetbuilder_b200.engine import Engine

    natoms, frames = 100_000, 1000
    sysm = synth.make_system(natoms, 0.02, seed=synth.BASE_SEED + 2)
    eng = Engine(local, sysm.atomname)
    pos = synth.frame_positions_device(sysm, frames, dev)
    types = eng.to_device_types(sysm.types)
    cell = sysm.cell()
    out = {"ell": torch.empty((frames, natoms, eng.max_bonds), dtype=torch.int32, device=dev),
           "keys": torch.empty((frames, natoms), dtype=torch.int64, device=dev),
           "labels": torch.empty((frames, natoms), dtype=torch.int32, device=dev)}
    for _ in range(3):
        eng.analyze(pos, cell, types, True, out=out)
    eng.check()
    eng.profile_read()
    eng.profile(True)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 5
    ev0.record()
    for _ in range(steps):
        eng.analyze(pos, cell, types, True, out=out)
    ev1.record()
    torch.cuda.synchronize()
    ms = max_over_ranks(ev0.elapsed_time(ev1)) / steps
    eng.profile(False)
    prof = eng.profile_read()
    nnz = int((out["ell"][:8] >= 0).sum().item())
    dbar = nnz / float(8 * natoms)
    alg = 24.0 + 4.0 + 4.0 * dbar
    kern = {k: {"launches": c, "ms_per_step": m / steps} for k, (c, m) in prof.items()}
    kb = kern.get("k_bonds") or kern.get("k_bonds_tile") or {}
    res = {"workload": "configs[1]: synthetic 100k-atom C/H/O ReaxFF dump, 1,000 frames, dump-only distance-based bond detection",
           "atom_frames_per_s": world * frames * natoms / (ms / 1e3), "ms_per_pass": ms, "mean_degree": dbar,
           "algorithmic_bytes_per_atom_frame": alg, "kernels": kern,
           "path_hbm_frac": alg * frames * natoms / (sum(v["ms_per_step"] for v in kern.values()) / 1e3) / 1e9 / peaks["hbm_gbs"]}
    if kb:
        res["k_bonds_hbm_frac"] = alg * frames * natoms / (kb["ms_per_step"] / 1e3) / 1e9 / peaks["hbm_gbs"]
    del pos, out
    eng.close()
    torch.cuda.empty_cache()
    return res


# --------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=FRAMES, help="frames per rank (testing; the workload is 1000)")
    ap.add_argument("--atoms", type=int, default=NATOMS, help="atoms per frame (testing; the workload is 1e6)")
    ap.add_argument("--clusters", type=int, default=KCLUST)
    ap.add_argument("--batch-frames", type=int, default=0, help="frames per device batch (0: ~8M atom-frames)")
    ap.add_argument("--pool-rows", type=int, default=1 << 21)
    ap.add_argument("--host-slots", type=int, default=64)
    ap.add_argument("--exchange", default="replicated", choices=["replicated", "allreduce"])
    ap.add_argument("--train-threads", type=int, default=1)
    ap.add_argument("--cache-gb", type=float, default=-1.0,
                    help="HBM for pass-B spectra (-1: what is free minus a reserve); small values force the host tier in short test runs")
    ap.add_argument("--host-cache-gb", type=float, default=-1.0,
                    help="host memory for pass-B spectra beyond the HBM tier (-1: what is available minus a reserve, 0: off)")
    ap.add_argument("--overlap-seeding", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-stages", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    from mddatasetbuilder_b200 import synth
    from mddatasetbuilder_b200.engine import Engine
    from mddatasetbuilder_b200.streaming import ResidentSource, StreamingSelector

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"  # keep stdout to the single JSON line
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local])

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    peaks, peak_kind = measured_peaks()
    natoms = args.atoms
    steps = max(1, args.steps)
    fps = max(1, args.frames // steps)             # frames per step
    frames = fps * steps                           # frames of the timed job (per rank)
    wframes = fps * max(args.warmup, 0)            # frames of the warm-up job
    bf = args.batch_frames or max(1, min(fps, (8 << 20) // natoms))
    sysm = synth.make_system(natoms, DENSITY, seed=synth.BASE_SEED + 3 + 1000 * rank)
    eng = Engine(local, sysm.atomname)
    nres = max(frames, wframes)
    pos = synth.frame_positions_device(sysm, nres, dev)
    rowptr, col, bo = synth.bond_table_device(sysm, nres, dev)
    cell = sysm.cell()
    kw = {"exchange": args.exchange} if world > 1 else {}

    def job(source, nfr):
        sel = StreamingSelector(eng, sysm.types, cutoff=5.0, n_clusters=args.clusters, n_each=1, seed=synth.BASE_SEED,
                                periodic=True, pool_rows=args.pool_rows, want_molecules=True, kmeans_kwargs=kw,
                                train_threads=args.train_threads, overlap_seeding=args.overlap_seeding,
                                cache_gb=None if args.cache_gb < 0 else args.cache_gb,
                                host_cache_gb=None if args.host_cache_gb < 0 else args.host_cache_gb)
        res = sel.run(source)
        return sel, res

    def resident(nfr):
        return ResidentSource(pos[:nfr], cell, bf, rowptr=rowptr[:nfr], col=col[:nfr], bo=bo[:nfr], rank=rank, world=world,
                              local=True)

    # ---- warm-up: a job of W steps -------------------------------------------------------------
    if wframes:
        job(resident(wframes), wframes)
    eng.check()
    eng.profile_read()

    # ---- timed: the job of K steps, inputs resident in HBM ---------------------------------------
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    eng.profile(True)
    barrier()
    torch.cuda.synchronize()
    l0 = eng.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.time()
    ev0.record()
    sel, res = job(resident(frames), frames)
    ev1.record()
    torch.cuda.synchronize()
    w1 = time.time()
    barrier()
    launches = eng.launches - l0 + sum(w.launches for w in sel.worker_engines())
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    eng.profile(False)
    prof = eng.profile_read()
    for w in sel.worker_engines():                 # the training threads' own contexts
        w.profile(False)
        for name, (cnt, ms) in w.profile_read().items():
            c0, m0 = prof.get(name, (0, 0.0))
            prof[name] = (c0 + cnt, m0 + ms)
    clocks = sampler.stop(w0, w1)
    ms_per_step = ms_total / steps
    units = world * frames * natoms
    value = units / (ms_total / 1e3)
    nchosen = sum(len(v) for v in res.chosen.values())
    phases = {k: max_over_ranks(v) for k, v in res.timings.items()}

    # ---- kernels: time share of every kernel of the job; roofline of the dominant one --------------
    kern = {k: {"launches": c, "ms": m} for k, (c, m) in prof.items()}
    gpu_ms = sum(v["ms"] for v in kern.values()) or float("nan")
    for v in kern.values():
        v["share"] = v["ms"] / gpu_ms
    # descriptor kernels by family
    desc_ms = sum(v["ms"] for k, v in kern.items() if k.startswith("k_tridiag") or k.startswith("k_tql") or k.startswith("k_eig"))
    # mean n and n^3 of the cutoff spheres (one batch of frames): useful flops of the tridiagonalisation
    cnt, _ = eng.compositions(pos[:1], cell, eng.to_device_types(sysm.types), 5.0, periodic=True)
    nn = cnt.sum(dim=1).double()
    n_mean, n3_mean = float(nn.mean().item()), float((nn ** 3).mean().item())
    del cnt, nn
    sc = res.stats.get("spectra_cache") or {"batches": 0, "of": 1}
    # pass B computes every descriptor, pass C only those of the batches whose spectra were not kept
    ntarget_passes = frames * natoms * (2.0 - (sc["batches"] + sc.get("host_batches", 0)) / max(sc["of"], 1))
    useful_flops = (8.0 / 3.0) * n3_mean * ntarget_passes
    fp64_peak = eng.fp64_peak_tflops()
    Dmean = float(np.mean([st["D"] for st in res.stats["fit"].values()])) if res.stats.get("fit") else 0.0
    dom = max((k for k in kern if k != "memset"), key=lambda k: kern[k]["ms"])
    desc_alg_bytes = (16.0 + 8.0 * Dmean) * ntarget_passes
    roofline = {"bound": "fp64", "kernel": "k_tridiag_* + k_tql (descriptor: Householder tridiagonalisation + QL)",
                "dominant_single_kernel": dom,
                "achieved": useful_flops / (desc_ms / 1e3) / 1e12 if desc_ms else None, "peak": fp64_peak, "unit": "TFLOP/s",
                "frac": (useful_flops / (desc_ms / 1e3) / 1e12 / fp64_peak) if desc_ms and fp64_peak else None,
                "traffic": None, "peak_source": "DFMA peak measured in this run (k_fp64_peak)",
                "useful_flops_per_target": (8.0 / 3.0) * n3_mean, "mean_sphere_atoms": n_mean,
                "descriptor_ms": desc_ms, "descriptor_ns_per_target": desc_ms * 1e6 / ntarget_passes if desc_ms else None,
                "hbm": {"algorithmic_bytes_per_target": 16.0 + 8.0 * Dmean,
                        "achieved_gbs": desc_alg_bytes / (desc_ms / 1e3) / 1e9 if desc_ms else None,
                        "peak_gbs": peaks["hbm_gbs"], "peak_source": peak_kind,
                        "frac": (desc_alg_bytes / (desc_ms / 1e3) / 1e9 / peaks["hbm_gbs"]) if desc_ms else None},
                "job_hbm": {"algorithmic_bytes_per_atom_frame": 727.0,
                            "achieved_gbs": 727.0 * frames * natoms / (ms_total / 1e3) / 1e9,
                            "frac": 727.0 * frames * natoms / (ms_total / 1e3) / 1e9 / peaks["hbm_gbs"],
                            "note": "SURVEY.md 8(d): keys 29 + connectivity 16 + composition 22 + descriptor 336 + assign 324 B"}}
    comm = None
    if world > 1:
        comm = {"exchange": args.exchange, "training_allreduce": res.stats.get("fit_comm"), "other": res.stats.get("comm"),
                "train_phase_s": phases.get("train"), "job_s": ms_total / 1e3}

    # ---- e2e: the same job from page-locked host arrays ------------------------------------------
    e2e = None
    if not args.no_e2e:
        hs = HostSource(torch, dev, pos, rowptr, col, bo, cell, frames, bf, rank, world, args.host_slots)
        del pos, rowptr, col, bo
        torch.cuda.empty_cache()
        barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        _, eres = job(hs, frames)
        nsel = sum(len(v) for v in eres.chosen.values())   # the selection is host data on return
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        barrier()
        e_ms = max_over_ranks((t1 - t0) * 1e3)
        same = eres.chosen == res.chosen if hs.slots >= frames else None
        e2e = {"value": units / (e_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": int(hs.h2d_bytes // steps),
               "d2h_bytes_per_step": int(nsel * 16 // steps + 1), "ms_per_step": e_ms / steps,
               "selected_structures": nsel, "same_selection_as_resident_run": same,
               "host_frames_kept": hs.slots,
               "api": "StreamingSelector.run(source) with page-locked host arrays (positions, CSR bond table): every "
                      "batch of each of the three passes is copied host -> device inside the timed region; the chosen "
                      "(step, atom) lists come back to the host",
               "phases_s": {k: max_over_ranks(v) for k, v in eres.timings.items()}}
        del hs

    stages = None
    if not args.no_stages:
        torch.cuda.empty_cache()
        stages = {"c2_bonds": measure_c2_bonds(torch, dev, local, max_over_ranks, world, peaks)}

    base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        base = cpu_baseline(natoms=natoms, sysm=sysm if natoms == NATOMS else None)

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": C3_TEXT if world == 1 else C4_TEXT, "atoms_per_frame": natoms,
                           "frames_per_rank": frames, "frames_total": frames * world, "frames_per_step": fps,
                           "frames_per_device_batch": bf, "density_per_A3": DENSITY, "n_clusters": args.clusters,
                           "pool_rows": args.pool_rows, "cutoff_A": 5.0,
                           "stages": "bond-type keys from the bond table + device group-by + sphere compositions + molecule "
                                     "labels (pass A); descriptors -> scaler bounds + training pool (pass B); MiniBatchKMeans "
                                     "per bond type; descriptors -> scale -> tcgen05 assignment (pass C); selection",
                           "step": "1/steps of the job's frames through all three passes; training + selection once per "
                                   "job, inside the timed region; warm-up = a job of `warmup` steps",
                           "bond_types": len(res.keys), "clustered_types": len(res.maxcount),
                           "selected_structures": nchosen, "spectra_cache": res.stats.get("spectra_cache"),
                           "hbm_peak_allocated_gb": res.stats.get("hbm_peak_allocated_gb"),
                           "allocator": res.stats.get("allocator"),
                           "parallelism": f"frame batches interleaved over {world} rank(s); training {args.exchange}, "
                                          f"bond types dealt over the ranks, {args.train_threads} concurrent fits per rank",
                           "l2": f"inputs are {frames * natoms * 45 / 1e9:.1f} GB per rank (> 126 MB L2)",
                           "seed": synth.BASE_SEED + 3},
                "phases_s": phases, "roofline": roofline, "kernels": kern, "fit": {str(k): v for k, v in res.stats["fit"].items()},
                "stages": stages, "cpu_baseline": base, "e2e": e2e, "comm": comm,
                "gpu_launches": int(launches), "clocks": clocks}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
