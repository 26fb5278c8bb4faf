"""GPU parity of the streaming passes (mddatasetbuilder_b200/streaming.py, csrc/stream.cu): the device
group-by against numpy, keys / molecule labels from a CSR bond table against the oracle, per-type
max_counter, the label histogram / pick kernels against numpy, and the streamed selection against the
in-memory `clusterdatas` path bit for bit (reference datasetbuilder.py:185-281, 351-384)."""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _t(engine, a, dtype=None):
    import torch

    return torch.as_tensor(np.ascontiguousarray(a), device=engine.device, dtype=dtype)


def _system(natoms, density=0.05, seed=11):
    from mddatasetbuilder_b200 import synth

    return synth, synth.make_system(natoms, density, seed=synth.BASE_SEED + seed)


@pytest.mark.parametrize("F,N,keep", [(3, 5000, False), (2, 20011, True), (1, 17, False)])
def test_group_keys_matches_numpy(engine, F, N, keep):
    rng = np.random.default_rng(F * N)
    pool = rng.integers(1, 2 ** 62, size=37, dtype=np.int64)
    keys = pool[rng.integers(0, len(pool), size=(F, N))]
    keys[0, : min(N, 40)] = pool[3]                      # a whole warp of one key
    kp = (rng.random((F, N)) < 0.7).astype(np.uint8) if keep else None
    table = engine.key_table(256)
    codes, order, seg = engine.group_keys(_t(engine, keys), table, step0=5, keep=None if kp is None else _t(engine, kp))
    # a second batch into the same table keeps the slots of known keys
    codes2, _, _ = engine.group_keys(_t(engine, keys[:1]), table, step0=5 + F, keep=None, want_order=False)
    engine.check()
    tk = table[0].cpu().numpy()
    tc = table[1].cpu().numpy()
    tf = table[2].cpu().numpy().view(np.uint64)
    codes = codes.cpu().numpy().view(np.uint16)
    order = order.cpu().numpy()
    seg = seg.cpu().numpy()
    flat = keys.reshape(-1)
    live = np.ones(F * N, bool) if kp is None else kp.reshape(-1).astype(bool)
    assert np.all(codes.reshape(-1)[~live] == 0xFFFF)
    assert np.array_equal(tk[codes.reshape(-1)[live]], flat[live]), "codes do not translate back to the keys"
    assert np.array_equal(tk[codes2.cpu().numpy().view(np.uint16)[0]], keys[0])
    assert seg[-1] == live.sum()
    for s in np.nonzero(tk != -1)[0]:
        want = np.nonzero(live & (flat == tk[s]))[0]
        got = order[seg[s]:seg[s + 1]]
        assert np.array_equal(got, want), "rows of a key are not the ascending atom-frames"
        n2 = int((keys[0] == tk[s]).sum())
        assert tc[s] == len(want) + n2
        if len(want):
            a = want[0]
            assert tf[s] == ((5 + a // N) << 32 | (a % N))


def test_keys_and_labels_from_csr(engine):
    import torch

    synth, s = _system(30000)
    F = 3
    rowptr, col, bo = synth.bond_table_device(s, F, engine.device)
    types = engine.to_device_types(s.types)
    keys = engine.bond_keys_csr(rowptr, bo, types).cpu().numpy().view(np.uint64)
    labels = engine.components_csr(rowptr, col).cpu().numpy()
    boh = bo.cpu().numpy()
    for f in range(F):
        assert np.array_equal(keys[f], O.bond_keys_bondfile(s.types, s.rowptr, boh[f])), "keys differ from the oracle"
        assert np.array_equal(labels[f], O.dps(s.rowptr, s.col)[2]), "molecule labels differ from dps"
    assert not np.array_equal(boh[0], boh[1])            # frames carry their own bond orders
    # the ELL route of the text parser gives the same keys
    width = 8
    nbr = np.full((s.natoms, width), -1, np.int32)
    bor = np.zeros((s.natoms, width))
    for i in range(s.natoms):
        a, b = s.rowptr[i], s.rowptr[i + 1]
        nbr[i, : b - a] = s.col[a:b]
        bor[i, : b - a] = boh[1][a:b]
    k2 = engine.bond_keys_bo(_t(engine, nbr[None]), _t(engine, bor[None]), types).cpu().numpy().view(np.uint64)
    assert np.array_equal(k2[0], keys[1])
    del torch


def test_type_maxcount(engine):
    rng = np.random.default_rng(3)
    T, nt, cap = 50000, 3, 64
    codes = rng.integers(0, 9, size=T).astype(np.uint16)
    codes[rng.random(T) < 0.05] = 0xFFFF
    codes[:64] = 4
    counts = rng.integers(0, 30, size=(T, nt)).astype(np.int32)
    import torch

    maxc = torch.zeros((cap, engine.MAXT), dtype=torch.int32, device=engine.device)
    engine.type_maxcount(_t(engine, codes.view(np.int16)), _t(engine, counts), maxc)
    m = maxc.cpu().numpy()
    for s in range(9):
        sel = codes == s
        assert np.array_equal(m[s, :nt], counts[sel].max(axis=0))
    assert m[9:].sum() == 0


def test_label_hist_and_pick_rows(engine):
    import torch

    rng = np.random.default_rng(9)
    rows, K = 300007, 1000
    labels = rng.integers(0, K, size=rows).astype(np.int32)
    labels[1000:6000] = 7                                # a dense run
    d = _t(engine, labels)
    hist = torch.zeros(K, dtype=torch.int32, device=engine.device)
    engine.label_hist(d, hist)
    assert np.array_equal(hist.cpu().numpy(), np.bincount(labels, minlength=K))
    q, want = [], []
    for _ in range(400):
        r0 = int(rng.integers(0, rows - 1))
        r1 = int(rng.integers(r0 + 1, rows + 1))
        k = int(rng.integers(0, K))
        idx = np.nonzero(labels[r0:r1] == k)[0]
        p = int(rng.integers(0, max(len(idx), 1) + 1))
        q.append((r0, r1, k, p))
        want.append(r0 + idx[p] if p < len(idx) else -1)
    q.append((0, rows, 7, 4999))
    want.append(np.nonzero(labels == 7)[0][4999])
    got = engine.pick_rows(d, np.array(q, dtype=np.int64)).cpu().numpy()
    assert np.array_equal(got, np.array(want))


def _in_memory_selection(engine, s, pos, bo_frames, K, n_each, seed, pool_rows):
    """The round-1 flow: every row of a type in memory -> scaler -> MiniBatchKMeans -> choose_per_cluster
    (with the training restricted to the selector's seeded pool when the type has more rows than pool_rows)."""
    import torch

    from mddatasetbuilder_b200.kmeans import MiniBatchKMeansB200, MinMaxScalerB200, choose_per_cluster

    F, N = pos.shape[0], pos.shape[1]
    types = engine.to_device_types(s.types)
    keys = np.stack([O.bond_keys_bondfile(s.types, s.rowptr, bo_frames[f]) for f in range(F)])
    first = {}
    for f in range(F):
        for i, k in enumerate(keys[f]):
            first.setdefault(int(k), (f, i))
    out = {}
    for k in sorted(first, key=lambda x: first[x]):
        ff, ii = np.nonzero(keys == np.uint64(k))
        stepatom = np.stack([ff, ii + 1], axis=1)
        if len(ff) <= K:
            out[k] = [(int(a), int(b)) for a, b in stepatom]
            continue
        targets = (ff * N + ii).astype(np.int32)
        counts, maxc = engine.compositions(pos, s.cell(), types, 5.0, targets=targets)
        X = engine.descriptors(pos, s.cell(), types, 5.0, counts, maxc, targets=targets)["X"]
        engine.check()
        MinMaxScalerB200(engine).fit_transform(X)
        n = len(ff)
        if n > pool_rows:
            rng = np.random.default_rng([seed & 0xFFFFFFFF, k & 0xFFFFFFFF, k >> 32])
            pidx = np.sort(rng.choice(n, size=pool_rows, replace=False, shuffle=False))
            pool = X[torch.as_tensor(pidx, device=engine.device)].contiguous()
        else:
            pool = X
        km = MiniBatchKMeansB200(engine, n_clusters=K, init_size=min(3 * K, int(pool.shape[0])), n_init=3, seed=seed)
        km.fit(pool, compute_labels=False)
        labels = km.predict(X).cpu().numpy()
        picks = choose_per_cluster(labels, K, n_each, seed)
        out[k] = [(int(stepatom[p, 0]), int(stepatom[p, 1])) for p in picks]
    return out


@pytest.mark.parametrize("pool_rows,batch,cache_gb,host_gb", [(1 << 21, 2, None, 0.0), (90, 3, 0.0, 0.0), (90, 2, 0.002, 0.0),
                                                             (90, 2, 0.0, 1.0), (1 << 21, 1, 0.001, 0.0008), (90, 2, 0.002, None)])
def test_streamed_selection_equals_in_memory(engine, pool_rows, batch, cache_gb, host_gb):
    """SURVEY.md hard part 3 / VERDICT item 2: the three-pass job picks exactly the rows the in-memory
    path picks -- all rows in the pool (first case) and with a seeded sub-sample as pool (second)."""
    from mddatasetbuilder_b200.streaming import ResidentSource, StreamingSelector

    synth, s = _system(1500, seed=21)
    F, K, n_each, seed = 7, 30, 2, 5
    pos = _t(engine, synth.frame_positions(s, F))
    rowptr, col, bo = synth.bond_table_device(s, F, engine.device)
    sel = StreamingSelector(engine, s.types, cutoff=5.0, n_clusters=K, n_each=n_each, seed=seed, periodic=True,
                            pool_rows=pool_rows, want_molecules=True, cache_gb=cache_gb, host_cache_gb=host_gb)
    res = sel.run(ResidentSource(pos, s.cell(), batch, rowptr=rowptr, col=col, bo=bo))
    nb = -(-F // batch)
    nc = res.stats["spectra_cache"]["batches"]       # pass C: spectra kept in HBM / in host memory / recomputed / a mix
    nh = res.stats["spectra_cache"]["host_batches"]
    assert (nc == nb) if cache_gb is None else (nc == 0) if cache_gb == 0.0 else (0 < nc < nb)
    if host_gb == 0.0:
        assert nh == 0
    elif host_gb == 0.0008:
        assert 0 < nh and 0 < nc and nc + nh < nb, (nc, nh, nb)    # all three routes in one job
    else:
        assert nc + nh == nb and nh > 0, (nc, nh, nb)
    want = _in_memory_selection(engine, s, pos, bo.cpu().numpy(), K, n_each, seed, sel.pool_rows)
    assert [int(k) for k in res.keys] == list(want.keys()), "bond types / first-seen order differ"
    assert len(sel.big) >= 3
    for k in want:
        got = res.chosen[k]
        if k in sel.big:
            assert got == want[k], f"type {O.key_string(k, s.atomname)}: streamed selection differs"
        else:
            assert sorted(got) == sorted(want[k])
    if pool_rows == 90:
        assert any(v["pool_rows"] < v["rows"] for v in res.stats["fit"].values()), "no type used a sub-sampled pool"


def test_streamed_dump_only_and_keep_mask(engine):
    """Dump-only keys (bonds from coordinates) and the model-deviation filter through the same job."""
    from mddatasetbuilder_b200.streaming import ResidentSource, StreamingSelector

    synth, s = _system(1200, density=0.04, seed=4)
    F = 4
    frames = synth.frame_positions(s, F)
    rng = np.random.default_rng(0)
    keep = (rng.random((F, s.natoms)) < 0.6).astype(np.uint8)
    sel = StreamingSelector(engine, s.types, n_clusters=10 ** 6, seed=1, periodic=True)
    res = sel.run(ResidentSource(_t(engine, frames), s.cell(), 3, keep=_t(engine, keep)))
    out = engine.analyze(frames, s.cell(), s.types, True)
    keys = out["keys"].cpu().numpy().view(np.uint64)
    for k in res.keys:
        ff, ii = np.nonzero((keys == np.uint64(k)) & (keep > 0))
        assert res.chosen[k] == sorted((int(a), int(b) + 1) for a, b in zip(ff, ii))
    assert sum(res.counts.values()) == int(keep.sum())


def test_two_rank_job_equals_single_rank():
    """C4: two ranks (NCCL) deal the batches of one trajectory between them and must select exactly what one
    rank selects (replicated training on the all-reduced pool); skipped on a single-GPU box."""
    import os
    import subprocess
    import sys

    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533",
                        os.path.join(root, "tools", "ddp_stream_check.py")], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "STREAM-DDP-OK" in r.stdout
