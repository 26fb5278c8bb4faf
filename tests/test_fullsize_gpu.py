"""Full-size frames (BASELINE.json configs[1]/[2] shapes: 100k and 1M atoms per frame): the O(N^2) oracle
cannot run there, so parity is checked through size-independent properties of the domain."""
import numpy as np
import pytest
import torch

from mddatasetbuilder_b200 import synth

pytestmark = pytest.mark.gpu
MAXBONDS = {"C": 4, "H": 1, "O": 2}


def _properties(s, out, nframes):
    ell = out["ell"].cpu().numpy()
    labels = out["labels"].cpu().numpy()
    keys = out["keys"].cpu().numpy().view(np.uint64)
    N = s.natoms
    maxb = np.array([MAXBONDS[s.atomname[t]] for t in s.types])
    for f in range(nframes):
        e = ell[f]
        valid = e >= 0
        deg = valid.sum(axis=1)
        # rows strictly ascending, padding only at the end, no self bonds
        assert np.all(valid[:, :-1] | ~valid[:, 1:])
        assert np.all(~(valid[:, :-1] & valid[:, 1:]) | (e[:, :-1] < e[:, 1:]))
        rows = np.repeat(np.arange(N), deg)
        cols = e[valid]
        assert np.all(rows != cols)
        # symmetry: the directed edge multiset equals its transpose
        a = np.sort(rows.astype(np.int64) * N + cols)
        b = np.sort(cols.astype(np.int64) * N + rows)
        assert np.array_equal(a, b), "bond list is not symmetric"
        # valence rule of the cleanup holds for every atom
        assert np.all(deg <= maxb)
        # labels: constant across every bond, equal to the smallest index of the component
        lab = labels[f]
        assert np.array_equal(lab[rows], lab[cols])
        assert np.all(lab <= np.arange(N))
        assert np.array_equal(lab[lab], lab), "label is not a fixed point (root of its own component)"
        mins = np.full(N, N, dtype=np.int64)
        np.minimum.at(mins, lab, np.arange(N))
        roots = np.unique(lab)
        assert np.array_equal(mins[roots], roots)
        # keys: element and degree packed as documented
        assert np.array_equal((keys[f] >> np.uint64(56)).astype(np.int64) - 1, s.types)
        assert np.array_equal(((keys[f] >> np.uint64(48)) & np.uint64(0xFF)).astype(np.int64), deg)
    return ell, labels, keys


@pytest.mark.parametrize("natoms,density,nframes", [(100_000, 0.02, 6), (1_000_000, 0.05, 2)])
def test_properties_and_chunking_idempotence(engine, natoms, density, nframes):
    s = synth.make_system(natoms, density, seed=synth.BASE_SEED + 2)
    engine.set_elements(s.atomname)
    pos = synth.frame_positions_device(s, nframes, engine.device)
    out = engine.analyze(pos, s.cell(), s.types)
    stats = engine.check()
    assert stats["overflow"] == 0
    ell, labels, keys = _properties(s, out, nframes)
    # frame 0 is the unperturbed generator geometry: away from the reactive contacts the detected bonds
    # are exactly the generator's topology, so the two edge sets agree on the vast majority of atoms
    deg0 = (ell[0] >= 0).sum(axis=1)
    same = deg0 == np.diff(s.rowptr)
    assert same.mean() > 0.97
    # idempotence under batching: one frame per launch gives the same answer as one launch for all
    engine.set_option("frames_per_launch", 1)
    try:
        out1 = engine.analyze(pos, s.cell(), s.types)
        engine.check()
    finally:
        engine.set_option("frames_per_launch", 0)
    assert torch.equal(out1["ell"], out["ell"]) and torch.equal(out1["labels"], out["labels"])
    assert torch.equal(out1["keys"], out["keys"])
    # host-buffer path returns the same bytes
    host = engine.analyze_host(pos[:2].cpu().numpy(), s.cell(), s.types)
    assert np.array_equal(host["ell"], ell[:2]) and np.array_equal(host["labels"], labels[:2])
    assert np.array_equal(host["keys"], keys[:2])
    # determinism: a checksum of per-frame checksums is stable across repeated runs
    c1 = int(out["ell"].long().sum().item()) ^ int(out["labels"].long().sum().item())
    out2 = engine.analyze(pos, s.cell(), s.types)
    c2 = int(out2["ell"].long().sum().item()) ^ int(out2["labels"].long().sum().item())
    assert c1 == c2
    engine.set_elements(["C", "H", "O"])


def test_descriptor_and_assignment_properties_100k(engine):
    """Descriptor rows of a 100k-atom frame: sorted, finite, permutation-invariant spectrum sums
    (trace identity: sum of eigenvalues = sum of the Coulomb diagonal = sum of Z^2.4/2 over the sphere)."""
    s = synth.make_system(100_000, 0.02, seed=synth.BASE_SEED + 2)
    engine.set_elements(s.atomname)
    pos = synth.frame_positions_device(s, 1, engine.device)
    types = engine.to_device_types(s.types)
    targets = np.arange(0, 100_000, 7, dtype=np.int32)
    counts, maxc = engine.compositions(pos, s.cell(), types, 5.0, targets=targets)
    out = engine.descriptors(pos, s.cell(), types, 5.0, counts, maxc, targets=targets, want_eig=True, eigcap=64)
    engine.check()
    X = out["X"].cpu().numpy()
    assert np.all(np.isfinite(X)) and np.all(np.diff(X, axis=1) >= 0)
    c = counts.cpu().numpy()
    n = out["n"].cpu().numpy()
    assert np.array_equal(c.sum(axis=1), n) and n.min() >= 1
    pads = np.array([{"C": 6, "H": 1, "O": 8}[a] ** 2.4 / 2 for a in s.atomname])
    eig = np.nan_to_num(out["eig"].cpu().numpy(), nan=0.0)
    trace = (c * pads).sum(axis=1)
    assert np.allclose(eig.sum(axis=1), trace, rtol=1e-9, atol=1e-9)
    # row = spectrum + pads: the row sum equals the padded trace
    assert np.allclose(X.sum(axis=1), (maxc * pads).sum(), rtol=1e-9)
    engine.set_elements(["C", "H", "O"])


# ---------------------------------------------------------------------------------------------
# Oracle comparisons AT the BASELINE config sizes: the cell-list oracle
# (oracle.c orc_connect_the_dots mode 1, O(N)) bit-compares whole frames, the O(N)-per-target descriptor
# oracle checks sampled targets, the perception restatement runs on full 100k-atom frames.
# ---------------------------------------------------------------------------------------------
def _Z(s):
    from oracle import oracle as O

    return np.array([O.ATOMIC_NUMBERS[s.atomname[t]] for t in s.types], dtype=np.int32)


@pytest.mark.parametrize("natoms,density,frames", [(100_000, 0.02, (0, 3)), (1_000_000, 0.05, (1,))])
def test_bonds_labels_keys_bit_exact_at_config_size(engine, natoms, density, frames):
    from oracle import oracle as O
    from tests.util import ell_to_canonical

    s = synth.make_system(natoms, density, seed=synth.BASE_SEED + 2)
    engine.set_elements(s.atomname)
    try:
        pos = synth.frame_positions_device(s, max(frames) + 1, engine.device)
        out = engine.analyze(pos, s.cell(), s.types)
        engine.check()
        Z = _Z(s)
        for f in frames:
            P = pos[f].cpu().numpy()
            bonds, _ = O.connect_the_dots(Z, P, s.cell(), True, mode=1)
            rp, col = O.bonds_to_csr(natoms, bonds)
            crp, ccol = O.canonical_csr(rp, col)
            grp, gcol = ell_to_canonical(out["ell"][f].cpu().numpy())
            assert np.array_equal(grp, crp) and np.array_equal(gcol, ccol), f"bond list of frame {f} differs from the oracle"
            assert np.array_equal(out["labels"][f].cpu().numpy(), O.dps(rp, col)[2]), "molecule labels differ"
            assert np.array_equal(out["keys"][f].cpu().numpy().view(np.uint64), O.bond_keys_degree(s.types, rp))
        # the bond-file route on the same system: keys and labels from the CSR table
        rowptr, colt, bo = synth.bond_table_device(s, 2, engine.device)
        keys = engine.bond_keys_csr(rowptr, bo, engine.to_device_types(s.types)).cpu().numpy().view(np.uint64)
        labels = engine.components_csr(rowptr, colt).cpu().numpy()
        boh = bo[1].cpu().numpy()
        assert np.array_equal(keys[1], O.bond_keys_bondfile(s.types, s.rowptr, boh))
        assert np.array_equal(labels[1], O.dps(s.rowptr, s.col)[2])
    finally:
        engine.set_elements(["C", "H", "O"])


@pytest.mark.parametrize("natoms,density", [(100_000, 0.02), (1_000_000, 0.05)])
def test_descriptor_rows_vs_oracle_at_config_size(engine, natoms, density):
    """2,000 sampled targets of one full-size frame: sphere membership bit-exact, padded sorted rows within
    |delta| <= 1e-5 * max(|ref|, 1e-3 * ||M||) of sphere cut + Coulomb matrix (oracle.c) + numpy eigh."""
    from concurrent.futures import ThreadPoolExecutor

    from oracle import oracle as O

    s = synth.make_system(natoms, density, seed=synth.BASE_SEED + 3)
    engine.set_elements(s.atomname)
    try:
        pos = synth.frame_positions_device(s, 2, engine.device)
        P = pos[1].cpu().numpy()
        types = engine.to_device_types(s.types)
        rng = np.random.default_rng(natoms)
        tg = np.sort(rng.choice(natoms, 2000, replace=False)).astype(np.int32)
        counts, maxc, members = engine.compositions(pos[1:2], s.cell(), types, 5.0, targets=tg, members_cap=96)
        out = engine.descriptors(pos[1:2], s.cell(), types, 5.0, counts, maxc, targets=tg, members=members)
        engine.check()
        X = out["X"].cpu().numpy()
        Z = _Z(s)
        diag = np.array([O.ATOMIC_NUMBERS[s.atomname[t]] ** 2.4 / 2 for t in s.types])
        chunks = np.array_split(tg, 32)
        with ThreadPoolExecutor(16) as ex:
            res = list(ex.map(lambda c: O.descriptor_gather(Z, diag, P, s.box, True, 5.0, c), chunks))
        mem = [m for _, ms, _ in res for m in ms]
        mats = [m for _, _, ms in res for m in ms]
        got_mem = members.cpu().numpy()
        for k in range(len(tg)):
            assert np.array_equal(np.sort(got_mem[k][got_mem[k] >= 0]), mem[k]), "sphere membership differs"
        ec = np.array([[int((s.types[m] == el).sum()) for el in range(len(s.atomname))] for m in mem])
        assert np.array_equal(ec, counts.cpu().numpy())
        pads = np.array([O.ATOMIC_NUMBERS[a] ** 2.4 / 2 for a in s.atomname])
        R, _ = O.feature_rows(O.coulomb_eigs(mats), ec, pads)
        mc = ec.max(axis=0)
        assert np.array_equal(mc, maxc)
        norms = np.array([np.abs(np.linalg.eigvalsh(M)).max() for M in mats])[:, None]
        tol = 1e-5 * np.maximum(np.abs(R), 1e-3 * norms)
        assert np.all(np.abs(X - R) <= tol), "descriptor rows outside the stated tolerance"
    finally:
        engine.set_elements(["C", "H", "O"])


def test_perception_full_frames_vs_restatement(engine):
    """Bond orders of two 100k-atom frames per density (tools/perceive_big_check.py as a test): every ELL slot
    equals the restated PerceiveBondOrders (oracle/perceive.py)."""
    from tests.test_perceive_gpu import ZNUM, _oracle_orders

    for density, dense in ((0.02, False), (0.10, True)):
        s = synth.make_system(100_000, density, seed=synth.BASE_SEED + 2, dense=dense)
        engine.set_elements(s.atomname)
        try:
            frames = synth.frame_positions(s, 2, sigma=0.08)
            cell = np.diag(s.box).reshape(9)
            out = engine.analyze(frames, cell, s.types)
            engine.check()
            got = engine.perceive(frames, cell, s.types, out["ell"], labels=out["labels"])
            engine.check()
            Z = [ZNUM[s.atomname[t]] for t in s.types]
            orders = got["orders"].cpu().numpy()
            ell = out["ell"].cpu().numpy()
            nmult = 0
            for f in range(2):
                want, _ = _oracle_orders(ell[f], frames[f], Z, cell, True)
                assert np.array_equal(want, orders[f]), f"density {density} frame {f}: bond orders differ"
                nmult += int((want > 1).sum()) // 2
            assert nmult > 1000
        finally:
            engine.set_elements(["C", "H", "O"])
