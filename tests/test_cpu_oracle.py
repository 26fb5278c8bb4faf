"""CPU suite (-m "not gpu"): pins the oracle against the reference's own known answers and the
committed golden fixtures, and checks the host-side logic and the C-ABI surface.  No GPU compute."""
import ctypes
import hashlib
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from mddatasetbuilder_b200 import synth
from oracle import oracle as O
from oracle import refrun

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(HERE, "golden")


@pytest.fixture(scope="module")
def golden():
    with open(os.path.join(GOLD, "golden.json")) as f:
        return json.load(f)


# ---------------------------------------------------------------- reference known answers
def test_reference_test_bond_vectors():
    # /root/reference/tests/test_bond.py:9-34
    cell = np.diag([20.0, 20.0, 20.0])
    pos = [[0.0, 0.0, 0.0], [19.0, 19.0, 19.0]]
    assert O.crd2bond([8, 8], pos, cell, True, readlevel=False) == [[1], [0]]
    assert O.crd2bond([8, 8], pos, cell, True, readlevel=True) == [[1], [1]]
    assert O.crd2bond([8, 8], pos, cell, False, readlevel=False) == [[], []]
    assert O.crd2bond([8, 8], pos, cell, False, readlevel=True) == [[], []]
    for mode in (0, 1):
        assert O.crd2bond([8, 8], pos, cell, True, mode=mode) == [[1], [0]]


def test_dps_known_answers_and_reference_extension():
    # SURVEY.md ground facts (probed with the reference's compiled dps)
    assert O.dps_lists([[1], [0, 2], [1], [], [5], [4]]) == [[0, 1, 2], [3], [4, 5]]
    assert O.dps_lists([[2, 1], [0], [0, 3], [2]]) == [[0, 1, 2, 3]]
    ref = O.ref_dps()
    if ref is None:
        pytest.skip("oracle/_ref not built (run make -C oracle where /root/reference exists)")
    rng = np.random.default_rng(0)
    for n in (1, 7, 60, 500):
        adj = [[] for _ in range(n)]
        for _ in range(int(n * 0.8)):
            a, b = rng.integers(0, n, 2)
            if a != b and b not in adj[a]:
                adj[a].append(int(b))
                adj[b].append(int(a))
        for row in adj:
            rng.shuffle(row)
        assert O.dps_lists(adj) == ref(adj)


def test_oracle_modes_agree_and_cleanup_is_exercised():
    s = synth.make_system(2500, 0.06, seed=synth.BASE_SEED + 77, reactive_fraction=0.05)
    Z = np.array([O.ATOMIC_NUMBERS[s.atomname[t]] for t in s.types])
    fr = synth.frame_positions(s, 2)[1]
    b0, nd0 = O.connect_the_dots(Z, fr, s.cell(), True, mode=0)
    b1, nd1 = O.connect_the_dots(Z, fr, s.cell(), True, mode=1)
    assert np.array_equal(b0, b1) and nd0 == nd1 and nd0 > 0
    b0, _ = O.connect_the_dots(Z, fr, s.cell(), False, mode=0)
    b1, _ = O.connect_the_dots(Z, fr, s.cell(), False, mode=1)
    assert np.array_equal(b0, b1)


# ---------------------------------------------------------------- golden fixtures
def _reader(kind, name):
    from mddatasetbuilder_b200.reader import TextReader

    return TextReader(kind, os.path.join(GOLD, "traj", name), nthreads=2)


def test_text_reader_matches_reference_readN_and_readcrd(golden):
    rd = _reader(0, "dump.small")
    assert rd.natoms == golden["natoms"] and rd.steplinenum == golden["dump_steplinenum"]
    assert rd.atomtype.tolist() == golden["atomtype"]
    pos, cell = rd.read(100, 1)
    assert len(pos) == golden["nframes"]
    for k, fr in enumerate(golden["frames"]):
        assert hashlib.sha256(np.ascontiguousarray(pos[k]).tobytes()).hexdigest() == fr["pos_sha"]
        assert np.array_equal(cell[k].reshape(3, 3), np.array(fr["cell"]))
    # stride and skip follow islice(zip_longest(...), 0, None, stride)
    rd.rewind()
    p2, _ = rd.read(100, 2)
    assert len(p2) == 3 and np.array_equal(p2[1], pos[2])
    rd.rewind()
    assert rd.skip(1, 2) == 1
    p3, _ = rd.read(1, 2)
    assert np.array_equal(p3[0], pos[2])
    # per-frame parse of a tuple of lines (with the None padding zip_longest produces)
    with open(os.path.join(GOLD, "traj", "dump.small")) as f:
        lines = f.readlines()
    n = rd.steplinenum
    pp, cc = rd.parse_lines(tuple(lines[n:2 * n]) + (None,))
    assert np.array_equal(pp, pos[1])


def test_bond_reader_keys_and_molecules_match_reference(golden):
    rb = _reader(1, "bonds.small")
    assert rb.natoms == golden["natoms"] and rb.steplinenum == golden["bond_steplinenum"]
    nbr, bo = rb.read(100, 1, width=8)
    names = golden["atomname"]
    types = np.array(golden["atomtype"]) - 1
    for k, fr in enumerate(golden["frames"]):
        adj = [[int(x) for x in row[row >= 0]] for row in nbr[k]]
        assert O.dps_lists(adj) == fr["bond_molecules"]
        rowptr, _ = O.lists_to_csr(adj)
        flat = np.concatenate([bo[k][i, : len(adj[i])] for i in range(len(adj))])
        keys = O.bond_keys_bondfile(types, rowptr, flat)
        got = {}
        for i, key in enumerate(keys):
            e, lv = O.unpack_key(key)
            got.setdefault((names[e], tuple(lv)), []).append(i + 1)
        want = {(n, tuple(lv)): ids for n, lv, ids in fr["bond_keys"]}
        assert got == want


def test_oracle_bonds_and_molecules_match_reference_through_shims(golden):
    rd = _reader(0, "dump.small")
    pos, cell = rd.read(100, 1)
    Z = np.array([O.ATOMIC_NUMBERS[golden["atomname"][t - 1]] for t in golden["atomtype"]])
    for k, fr in enumerate(golden["frames"]):
        for mode in (0, 1):
            adj = O.crd2bond(Z, pos[k], cell[k].reshape(3, 3), True, mode=mode)
            assert adj == fr["dump_adjacency"]
        assert O.dps_lists(adj) == fr["dump_molecules"]
        # keys: the adjacency above (Open Babel's bond order) through the PerceiveBondOrders restatement
        from tests.test_perceive_gpu import _uc
        from oracle.perceive import perceive_bond_orders

        uc = _uc(cell[k])
        P = pos[k]
        orders, _ = perceive_bond_orders([int(z) for z in Z], adj,
                                         lambda a, b: tuple(float(x) for x in uc.mic((P[a] - P[b]).reshape(1, 3))[0]))
        want = {(n, tuple(lv)): ids for n, lv, ids in fr["dump_keys"]}
        got = {}
        for i, a in enumerate(adj):
            got.setdefault((golden["atomname"][golden["atomtype"][i] - 1], tuple(sorted(orders[i]))), []).append(i + 1)
        assert got == want


def test_oracle_descriptors_match_reference_writestepmatrix(golden):
    rd = _reader(0, "dump.small")
    pos, cell = rd.read(100, 1)
    names = golden["atomname"]
    types = np.array(golden["atomtype"]) - 1
    Z = np.array([O.ATOMIC_NUMBERS[names[t]] for t in types])
    diag = O.coulomb_diag(names)
    dba = np.array([diag[names[t]] for t in types])
    L = np.diag(cell[1].reshape(3, 3))
    rows = golden["stepmatrix_frame1"]
    targets = np.array([r["atom"] - 1 for r in rows], dtype=np.int32)
    counts, members, mats = O.descriptor_gather(Z, dba, pos[1], L, True, 5.0, targets)
    eigs = O.coulomb_eigs(mats)
    for r, m, e in zip(rows, members, eigs):
        assert {names[t]: int(c) for t, c in enumerate(np.bincount(types[m], minlength=len(names))) if c} == r["counter"]
        assert np.allclose(e, r["eig"], rtol=1e-10, atol=1e-10)


@pytest.mark.skipif(not refrun.available(), reason="reference tree not mounted")
def test_golden_fixtures_are_reproducible_from_the_reference(golden, tmp_path):
    """Re-run a slice of make_golden.py against /root/reference and compare with the committed file."""
    refrun.load()
    from mddatasetbuilder.detect import DetectDump

    dd = DetectDump(filename=os.path.join(GOLD, "traj", "dump.small"), atomname=np.array(golden["atomname"]), pbc=True)
    with open(os.path.join(GOLD, "traj", "dump.small")) as f:
        lines = f.readlines()
    n = dd.steplinenum
    mols, _ = dd.readmolecule(tuple(lines[2 * n:3 * n]))
    assert [[int(a) for a in m] for m in mols] == golden["frames"][2]["dump_molecules"]


# ---------------------------------------------------------------- C ABI surface
def test_c_abi_exports_every_declared_symbol():
    from mddatasetbuilder_b200 import _lib

    with open(os.path.join(ROOT, "include", "mdb_b200.h")) as f:
        hdr = f.read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(mdb_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 30
    L = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} is declared in include/mdb_b200.h but not exported"
    assert declared == set(_lib.SIGNATURES), "ctypes prototypes out of sync with the header"
    assert _lib.lib().mdb_version() >= 100


def test_cli_help_runs_without_gpu():
    # reference tests/test_cli.py:7-9
    subprocess.check_output([sys.executable, "-m", "mddatasetbuilder_b200", "-h"], cwd=ROOT)


# ---------------------------------------------------------------- host logic
def test_writer_formats_and_wrap():
    from mddatasetbuilder_b200.writer import detect_multiplicity, wrap_positions

    assert detect_multiplicity(["O", "O"]) == 3
    assert detect_multiplicity(["H"]) == 2 and detect_multiplicity(["C", "H", "H", "H", "H"]) == 1
    sys.path.insert(0, os.path.join(ROOT, "oracle", "shims"))
    try:
        from ase import Atoms
    finally:
        sys.path.pop(0)
    rng = np.random.default_rng(2)
    cell = np.diag([11.0, 12.5, 9.0])
    pos = rng.uniform(-5, 20, (30, 3))
    a = Atoms("H30", positions=pos.copy(), cell=cell, pbc=True)
    center = pos[3] / np.array([11.0, 12.5, 9.0])
    a.wrap(center=center, pbc=a.get_pbc())
    assert np.array_equal(wrap_positions(pos, cell, center, True), a.positions)


def test_choose_per_cluster_single_and_sharded_agree():
    from mddatasetbuilder_b200.kmeans import choose_per_cluster

    rng = np.random.default_rng(4)
    labels = rng.integers(0, 40, 1000)
    labels[labels == 7] = 8  # an empty cluster
    picks, tags = choose_per_cluster(labels, 40, 2, seed=3, with_tags=True)
    assert len(picks) == 2 * len(np.unique(labels))
    assert all(labels[p] == t[0] for p, t in zip(picks, tags))
    # the same rows split over two ranks: the union of the ranks' picks is the single-process pick
    cut = 600
    parts = [labels[:cut], labels[cut:]]
    allc = np.stack([np.bincount(p, minlength=40) for p in parts])
    merged = {}
    for r, (p, off) in enumerate(zip(parts, (0, cut))):
        pk, tg = choose_per_cluster(p, 40, 2, seed=3, rank=r, world=2, all_counts=allc, with_tags=True)
        for i, t in zip(pk, tg):
            merged[tuple(t)] = labels[off + i]
    assert len(merged) == len(picks)
    assert all(merged[tuple(t)] == t[0] for t in tags)


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mddatasetbuilder_b200.kmeans import choose_per_cluster

    rng = np.random.default_rng(11)
    labels = rng.integers(0, 25, 800)
    mine = labels[rank::world]
    counts = torch.as_tensor(np.bincount(mine, minlength=25))
    allc = [torch.empty_like(counts) for _ in range(world)]
    dist.all_gather(allc, counts)
    picks, tags = choose_per_cluster(mine, 25, 1, seed=5, rank=rank, world=world,
                                     all_counts=torch.stack(allc).numpy(), with_tags=True)
    gathered = [None] * world
    dist.all_gather_object(gathered, [(int(t[0]), int(t[1]), int(mine[p])) for p, t in zip(picks, tags)])
    # scaler-bound reduction used by MinMaxScalerB200
    mn = torch.tensor([float(rank), 5.0 - rank])
    dist.all_reduce(mn, op=dist.ReduceOp.MIN)
    if rank == 0:
        q.put(([x for part in gathered for x in part], mn.tolist()))
    dist.destroy_process_group()


def test_distributed_selection_bookkeeping_gloo_world2():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    merged, mn = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert len(merged) == 25 and sorted(k for k, _, _ in merged) == list(range(25))
    assert all(k == lab for k, _, lab in merged), "every pick must belong to the cluster it was drawn for"
    assert mn == [0.0, 4.0]


def test_frame_boundary_search_parallel_equals_sequential(tmp_path):
    """The multi-threaded newline counting that slices frames (textio.cpp: advance_groups) delivers the
    same frames as the line-by-line walk, for every stride / batch size, with and without a final
    newline, and drops a truncated last frame (datasetbuilder.py:616-638 semantics)."""
    from mddatasetbuilder_b200 import synth
    from mddatasetbuilder_b200.reader import KIND_DUMP, TextReader

    s = synth.make_system(2500, 0.05, seed=synth.BASE_SEED + 71)
    base = str(tmp_path / "dump.full")
    synth.write_lammps_dump(base, s, synth.frame_positions(s, 9))
    data = open(base, "rb").read()

    def read_all(path, nt, stride, batch):
        r = TextReader(KIND_DUMP, path, nthreads=nt)
        out = []
        while True:
            pos, _ = r.read(batch, stride)
            if len(pos) == 0:
                break
            out.append(pos.copy())
        r.close()
        return np.concatenate(out)

    for name, blob, nframes in (("full", data, 9), ("nonl", data[:-1], 9), ("trunc", data[:-4000], 8)):
        path = str(tmp_path / ("dump." + name))
        open(path, "wb").write(blob)
        for stride in (1, 2, 4):
            for batch in (1, 3, 20):
                a = read_all(path, 1, stride, batch)
                b = read_all(path, 4, stride, batch)
                assert a.shape == b.shape == ((nframes + stride - 1) // stride, 2500, 3)
                assert np.array_equal(a, b)


def test_native_xyz_batch_writer_is_byte_identical(tmp_path):
    """mdb_write_xyz_batch writes the bytes of the reference's per-structure writer (ASE simple xyz,
    datasetbuilder.py:577-585; the Python restatement in writer.write_xyz is pinned by the golden files)."""
    from mddatasetbuilder_b200.writer import write_xyz, write_xyz_batch

    rng = np.random.default_rng(0)
    pa, pb, syms, poss = [], [], [], []
    for k in range(120):
        n = int(rng.integers(1, 50))
        s = rng.choice(np.array(["H", "C", "O", "Cl", "Na"]), n)
        p = rng.normal(0, 10.0 ** rng.integers(-3, 4), (n, 3))
        if k == 0:
            p[0] = [0.0, -0.0, 1e-17]
        if k == 1:
            p[0] = [123456789.123456789, -1e15, 5e-16]
        pa.append(str(tmp_path / f"a{k}.xyz"))
        pb.append(str(tmp_path / f"b{k}.xyz"))
        syms.append(s)
        poss.append(p)
        write_xyz(pa[-1], s, p)
    write_xyz_batch(pb, syms, poss, nthreads=3)
    for a, b in zip(pa, pb):
        assert open(a, "rb").read() == open(b, "rb").read()
    with pytest.raises(RuntimeError):
        write_xyz_batch([str(tmp_path / "no_such_dir" / "x.xyz")], syms[:1], poss[:1])


@pytest.mark.parametrize("fragment", [False, True])
@pytest.mark.parametrize("keywords", [["%nproc=4\n#force mn15/6-31g**"],
                                      ["%nproc=4\n#force mn15/6-31g**", "%nproc=4\n#force mn15/def2tzvp geom=chk", "%mem=2GB\n#sp"]])
def test_native_gjf_batch_writer_is_byte_identical(tmp_path, fragment, keywords):
    """mdb_write_gjf_batch writes the bytes of the reference's _convertgjf (datasetbuilder.py:468-521; the Python
    restatement in writer.write_gjf is pinned by the golden files): multiplicities incl. the O2 triplet, fragments,
    %chk / --link1-- sections, %.5f coordinates."""
    from mddatasetbuilder_b200.writer import write_gjf, write_gjf_batch

    rng = np.random.default_rng(1)
    pa, pb, syms, poss, sizes = [], [], [], [], []
    for k in range(150):
        nm = int(rng.integers(1, 6))
        ms = [int(rng.integers(1, 9)) for _ in range(nm)]
        s = []
        for q, m in enumerate(ms):
            if k % 7 == 0 and q == 0:
                ms[0] = 2
                s += ["O", "O"]                  # an O2 molecule: triplet
            else:
                s += list(rng.choice(np.array(["H", "C", "O", "N", "Cl"]), m))
        s = np.array(s)
        p = rng.normal(0, 10.0 ** rng.integers(-3, 3), (len(s), 3))
        if k == 1:
            p[0] = [0.0, -0.0, -4e-6]
        if k == 2:
            p[0] = [0.000005, 1.000015, -2.5e-6]  # halves at the fifth decimal
        idx, at = [], 0
        for m in ms:
            idx.append(range(at, at + m))
            at += m
        pa.append(str(tmp_path / f"a_{k}.x.gjf"))
        pb.append(str(tmp_path / f"b_{k}.x.gjf"))
        syms.append(s)
        poss.append(p)
        sizes.append(ms)
        write_gjf(pa[-1], idx, s, p, keywords, fragment)
    write_gjf_batch(pb, sizes, syms, poss, keywords, fragment, nthreads=3)
    for a, b in zip(pa, pb):
        # the %chk line carries the file's own name: compare with the names aligned
        assert open(a, "rb").read().replace(b"a_", b"b_") == open(b, "rb").read(), a
    with pytest.raises(RuntimeError):
        write_gjf_batch([str(tmp_path / "no_such_dir" / "x.gjf")], sizes[:1], syms[:1], poss[:1], keywords, fragment)


def test_frame_wrapper_equals_wrap_positions_bit_for_bit():
    """The per-frame wrapper used by the output cut runs the same NumPy operations as wrap_positions (ASE's
    Atoms.wrap as called at datasetbuilder.py:572-576, pinned by the golden xyz files): identical bytes for
    orthorhombic and tilted cells, full and partial periodicity."""
    from mddatasetbuilder_b200.writer import FrameWrapper, wrap_positions

    rng = np.random.default_rng(0)
    for t in range(600):
        cell = np.diag(rng.uniform(5, 60, 3)) if t % 3 else rng.uniform(-5, 40, (3, 3)) + np.diag([30, 30, 30])
        pbc = True if t % 4 else [bool(x) for x in rng.integers(0, 2, 3)]
        pos = rng.normal(0, 40, (200, 3))
        idx = rng.integers(0, 200, int(rng.integers(1, 60)))
        center = pos[idx[0]] / np.linalg.norm(cell, axis=1)
        a = wrap_positions(pos[idx], cell, center, pbc)
        b = FrameWrapper(cell, pbc)(pos[idx], center)
        assert a.tobytes() == b.tobytes()


def test_minibatch_draw_equals_numpy_choice():
    """kmeans.py draws its mini-batches like scikit-learn 1.9 (`RandomState.choice(n, B, p=uniform)` =
    searchsorted on the cdf); the fast form used there is the same function of the same variates."""
    for n in (5, 900, 12345, 1 << 18):
        p = np.ones(n) / n
        rs1, rs2 = np.random.RandomState(3), np.random.RandomState(3)
        cdf = np.cumsum(np.ones(n) / float(n))
        cdf /= cdf[-1]
        for _ in range(20):
            want = rs1.choice(n, 1024, p=p, replace=True)
            u = rs2.random_sample(1024)
            c = np.minimum((u * n).astype(np.int64), n - 1)
            for _ in range(3):
                c = c - ((c > 0) & (cdf[np.maximum(c - 1, 0)] > u))
                c = c + ((c <= n - 1) & (cdf[np.minimum(c, n - 1)] <= u))
            assert np.array_equal(want, c)


def test_host_memory_budget_helper():
    """The host tier of the spectra cache sizes itself from MemAvailable capped by the cgroup limit."""
    from mddatasetbuilder_b200.streaming import _host_memory_available

    v = _host_memory_available()
    assert isinstance(v, int) and 0 < v < (1 << 50)
