"""GPU parity: sphere composition (bit-exact) and Coulomb-matrix descriptors (1e-5 relative)
vs the CPU oracle (ASE MIC restated in oracle.c + the real numpy.linalg.eigh)."""
import numpy as np
import pytest

from mddatasetbuilder_b200 import synth
from oracle import oracle as O

pytestmark = pytest.mark.gpu

# tolerance stated by the task: descriptors within 1e-5 relative (floored at 1e-3 * ||M||_2)
RTOL = 1e-5


def _oracle(s, pos, targets, cutoff=5.0, periodic=True):
    Z = np.array([O.ATOMIC_NUMBERS[s.atomname[t]] for t in s.types], dtype=np.int32)
    diag = O.coulomb_diag(s.atomname)
    diag_by_atom = np.array([diag[s.atomname[t]] for t in s.types])
    counts_n, members, mats = O.descriptor_gather(Z, diag_by_atom, pos, s.box, periodic, cutoff, targets)
    eigs = O.coulomb_eigs(mats)
    comp = np.zeros((len(targets), len(s.atomname)), dtype=np.int64)
    for k, m in enumerate(members):
        comp[k] = np.bincount(s.types[m], minlength=len(s.atomname))
    norms = np.array([np.abs(e).max() for e in eigs])
    return counts_n, comp, eigs, norms


def _check(engine, s, pos, targets, cutoff=5.0, periodic=True):
    engine.set_elements(s.atomname)
    counts_n, comp, eigs, norms = _oracle(s, pos, targets, cutoff, periodic)
    counts, maxc = engine.compositions(pos, s.cell(), s.types, cutoff, targets=targets, periodic=periodic)
    assert np.array_equal(counts.cpu().numpy(), comp), "sphere composition differs"
    assert np.array_equal(maxc, comp.max(axis=0))
    out = engine.descriptors(pos, s.cell(), s.types, cutoff, counts, maxc, targets=targets, periodic=periodic,
                             want_eig=True, eigcap=int(counts_n.max()))
    engine.check()
    n = out["n"].cpu().numpy()
    assert np.array_equal(n, counts_n), "cluster sizes differ"
    eig = out["eig"].cpu().numpy()
    worst = 0.0
    for k, e in enumerate(eigs):
        got = eig[k, : len(e)]
        tol = RTOL * np.maximum(np.abs(e), 1e-3 * norms[k])
        err = np.abs(got - e)
        worst = max(worst, float((err / tol).max()))
        assert np.all(err <= tol), f"target {k}: eigenvalues off by {err.max():.3e}"
        assert np.all(np.isnan(eig[k, len(e):]))
    pads = np.array([O.coulomb_diag(s.atomname)[a] for a in s.atomname])
    Xref, maxref = O.feature_rows(eigs, comp, pads)
    X = out["X"].cpu().numpy()
    assert X.shape == Xref.shape
    tolX = RTOL * np.maximum(np.abs(Xref), 1e-3 * norms[:, None])
    assert np.all(np.abs(X - Xref) <= tolX)
    assert np.all(np.diff(X, axis=1) >= 0), "feature rows must be sorted ascending"
    return worst


@pytest.mark.parametrize("natoms,density,dense", [(3000, 0.02, False), (3000, 0.05, False), (2000, 0.11, True)])
def test_all_atoms(engine, natoms, density, dense):
    s = synth.make_system(natoms, density, seed=synth.BASE_SEED + 21 + natoms, dense=dense)
    pos = synth.frame_positions(s, 2)[1]
    worst = _check(engine, s, pos[None], np.arange(natoms, dtype=np.int32))
    assert worst < 1.0


def test_target_subset_two_frames(engine):
    s = synth.make_system(2500, 0.05, seed=synth.BASE_SEED + 31)
    frames = synth.frame_positions(s, 2)
    rng = np.random.default_rng(3)
    t1 = np.sort(rng.choice(2500, 300, replace=False)).astype(np.int32)
    engine.set_elements(s.atomname)
    targets = np.concatenate([t1, 2500 + t1[::2]]).astype(np.int32)
    counts, maxc = engine.compositions(frames, s.cell(), s.types, 5.0, targets=targets)
    out = engine.descriptors(frames, s.cell(), s.types, 5.0, counts, maxc, targets=targets, want_eig=True, eigcap=96)
    engine.check()
    for f, tt, sl in ((0, t1, slice(0, 300)), (1, t1[::2], slice(300, 450))):
        cn, comp, eigs, norms = _oracle(s, frames[f], tt)
        assert np.array_equal(counts.cpu().numpy()[sl], comp)
        eig = out["eig"].cpu().numpy()[sl]
        for k, e in enumerate(eigs):
            assert np.all(np.abs(eig[k, : len(e)] - e) <= RTOL * np.maximum(np.abs(e), 1e-3 * norms[k]))


def test_small_box_single_cell_axes(engine):
    # L < 3 * cutoff: every axis uses the single-cell minimum-image path
    s = synth.make_system(100, 0.05, seed=synth.BASE_SEED + 41)
    assert s.box[0] < 15.0
    _check(engine, s, s.pos0[None], np.arange(100, dtype=np.int32))


def test_small_dense_box_more_than_255_atoms_in_the_single_cell(engine):
    # 300 atoms in a 14.4 A box: one cell per axis under a 5 A cutoff holds the whole frame -- more than the
    # 8 rank bits of a normal frame index; frames with few cells get wider rank fields
    s = synth.make_system(300, 0.10, seed=synth.BASE_SEED + 43, dense=True)
    assert 10.0 < s.box[0] < 15.0
    _check(engine, s, s.pos0[None], np.arange(0, 300, 3, dtype=np.int32))


def test_cutoff_above_half_the_cell_is_refused(engine):
    s = synth.make_system(60, 0.05, seed=synth.BASE_SEED + 44)
    engine.set_elements(s.atomname)
    with pytest.raises(RuntimeError, match="half of a cell height"):
        engine.compositions(s.pos0[None], s.cell(), s.types, 0.51 * float(s.box[0]), targets=np.arange(4, dtype=np.int32))
    # an empty target list is no work, not "every atom"
    counts, maxc = engine.compositions(s.pos0[None], s.cell(), s.types, 4.0, targets=np.zeros(0, dtype=np.int32))
    engine.check()
    assert tuple(counts.shape) == (0, len(s.atomname)) and not np.any(maxc)


def test_large_clusters_every_size_class(engine):
    # dense phase with a 5.4 A cutoff: cluster sizes run past 96, so the register kernels (<= 64 atoms)
    # and both block-per-matrix classes (<= 96, <= 128) are all exercised
    s = synth.make_system(1500, 0.11, seed=synth.BASE_SEED + 47, dense=True)
    targets = np.arange(0, 1500, 2, dtype=np.int32)
    engine.set_elements(s.atomname)
    counts, _ = engine.compositions(s.pos0[None], s.cell(), s.types, 5.4, targets=targets)
    n = counts.sum(dim=1).cpu().numpy()
    assert n.max() > 96 and n.max() <= 128 and n.min() <= 64, (n.min(), n.max())
    worst = _check(engine, s, s.pos0[None], targets, cutoff=5.4)
    assert worst < 1.0


def test_clusters_up_to_160_atoms(engine):
    # 5.9 A in the dense phase: the largest shared-memory class (<= 160 atoms)
    s = synth.make_system(1500, 0.11, seed=synth.BASE_SEED + 47, dense=True)
    targets = np.arange(0, 1500, 5, dtype=np.int32)
    engine.set_elements(s.atomname)
    counts, maxc = engine.compositions(s.pos0[None], s.cell(), s.types, 5.9, targets=targets)
    n = counts.sum(dim=1).cpu().numpy()
    assert n.max() > 128 and n.max() <= 160, (n.min(), n.max())
    worst = _check(engine, s, s.pos0[None], targets, cutoff=5.9)
    assert worst < 1.0


@pytest.mark.parametrize("cutoff,lo,hi", [(8.0, 161, 320), (9.7, 330, 512)])
def test_clusters_beyond_160_atoms_global_memory_class(engine, cutoff, lo, hi):
    # larger cutoffs in the dense phase: 161..512 atoms per sphere go through the global-memory Householder class
    # (k_tridiag_glb + k_tql<512>), same tolerance against numpy.linalg.eigh as every other class
    s = synth.make_system(1500, 0.11, seed=synth.BASE_SEED + 47, dense=True)
    targets = np.arange(0, 1500, 25, dtype=np.int32)
    engine.set_elements(s.atomname)
    counts, maxc = engine.compositions(s.pos0[None], s.cell(), s.types, cutoff, targets=targets)
    n = counts.sum(dim=1).cpu().numpy()
    assert n.min() >= lo and n.max() <= hi, (n.min(), n.max())
    worst = _check(engine, s, s.pos0[None], targets, cutoff=cutoff)
    assert worst < 1.0


def test_mixed_small_and_large_clusters_in_one_call(engine):
    # 7 A in the dense phase: 133..197 atoms, the shared-memory classes and the global-memory class in one call
    s = synth.make_system(1500, 0.11, seed=synth.BASE_SEED + 47, dense=True)
    engine.set_elements(s.atomname)
    targets = np.arange(0, 1500, 30, dtype=np.int32)
    counts, maxc = engine.compositions(s.pos0[None], s.cell(), s.types, 7.0, targets=targets)
    n = counts.sum(dim=1).cpu().numpy()
    assert n.min() <= 160 < n.max(), (n.min(), n.max())
    worst = _check(engine, s, s.pos0[None], targets, cutoff=7.0)
    assert worst < 1.0


def test_more_than_512_atoms_in_a_sphere_is_refused(engine):
    s = synth.make_system(1500, 0.11, seed=synth.BASE_SEED + 47, dense=True)
    targets = np.arange(0, 1500, 5, dtype=np.int32)
    engine.set_elements(s.atomname)
    counts, maxc = engine.compositions(s.pos0[None], s.cell(), s.types, 11.5, targets=targets[:8])
    assert int(counts.sum(dim=1).max()) > 512
    with pytest.raises(RuntimeError, match="more than 512 atoms"):
        engine.descriptors(s.pos0[None], s.cell(), s.types, 11.5, counts, maxc, targets=targets[:8])


def test_descriptors_are_reproducible_bit_for_bit(engine):
    # the matrix is built in ascending atom index whatever order the cell scan delivered the members in,
    # so two runs (whose cell lists are filled in different atomic orders) give identical rows
    import torch

    s = synth.make_system(3000, 0.05, seed=synth.BASE_SEED + 49)
    engine.set_elements(s.atomname)
    pos = synth.frame_positions(s, 3)
    rows = []
    for _ in range(3):
        counts, maxc = engine.compositions(pos, s.cell(), s.types, 5.0)
        rows.append(engine.descriptors(pos, s.cell(), s.types, 5.0, counts, maxc)["X"].clone())
    assert torch.equal(rows[0], rows[1]) and torch.equal(rows[0], rows[2])


def test_nonperiodic_and_other_cutoff(engine):
    s = synth.make_system(1500, 0.05, seed=synth.BASE_SEED + 43)
    _check(engine, s, s.pos0[None], np.arange(0, 1500, 3, dtype=np.int32), cutoff=4.0, periodic=False)


def test_reference_feature_assembly_closed_form():
    """The closed form used on the GPU equals the reference's parent-side loop (datasetbuilder.py:236-275)."""
    from collections import Counter

    rng = np.random.default_rng(0)
    names = ["C", "H", "O"]
    diag = O.coulomb_diag(names)
    results, eigs, comp = [], [], []
    for _ in range(40):
        c = rng.integers(0, 5, 3)
        if c.sum() == 0:
            c[1] = 1
        e = np.sort(rng.normal(10, 20, int(c.sum())))
        cnt = Counter({n: int(k) for n, k in zip(names, c) if k})
        results.append((e, cnt))
        eigs.append(e)
        comp.append(c)
    ref, maxc = O.reference_feature_assembly(results, diag)
    X, mx = O.feature_rows(eigs, np.array(comp), np.array([diag[n] for n in names]))
    assert np.array_equal(ref, X)


def test_one_pass_path_members_and_feature_rows(engine):
    """compositions(members) -> descriptors(members, spectra only) -> feature_rows gives bit for bit the rows
    of the two-scan path, also when the final per-element maxima are larger than this batch's."""
    import torch

    s = synth.make_system(4000, 0.05, seed=synth.BASE_SEED + 53)
    engine.set_elements(s.atomname)
    pos = synth.frame_positions(s, 2)
    targets = np.arange(0, 2 * 4000, 3, dtype=np.int32)
    counts, maxc = engine.compositions(pos, s.cell(), s.types, 5.0, targets=targets)
    ref = engine.descriptors(pos, s.cell(), s.types, 5.0, counts, maxc, targets=targets, want_eig=True, eigcap=64)
    nmax = int(counts.sum(dim=1).max().item())
    counts2, maxc2, members = engine.compositions(pos, s.cell(), s.types, 5.0, targets=targets, members_cap=nmax)
    engine.check()
    assert torch.equal(counts, counts2) and np.array_equal(maxc, maxc2)
    m = members.cpu().numpy()
    assert np.array_equal((m >= 0).sum(axis=1), counts.sum(dim=1).cpu().numpy())
    out = engine.descriptors(pos, s.cell(), s.types, 5.0, counts, maxc, targets=targets, want_X=False, want_eig=True,
                             eigcap=nmax, members=members)
    engine.check()
    assert torch.equal(out["n"], ref["n"])
    assert torch.equal(out["eig"].nan_to_num(0.0), ref["eig"][:, :nmax].nan_to_num(0.0))
    X = engine.feature_rows(out["eig"], out["n"], counts, maxc)
    assert torch.equal(X, ref["X"])
    bigger = maxc + np.array([2, 0, 1], dtype=np.int32)
    ref2 = engine.descriptors(pos, s.cell(), s.types, 5.0, counts, bigger, targets=targets)
    assert torch.equal(engine.feature_rows(out["eig"], out["n"], counts, bigger), ref2["X"])
    # a list that is too short is reported
    with pytest.raises(RuntimeError):
        engine.compositions(pos, s.cell(), s.types, 5.0, targets=targets, members_cap=max(nmax - 3, 1))
        engine.check()


def _triclinic_oracle(numbers, pos, cell, targets, cutoff, diagz):
    """Sphere cut and Coulomb spectrum of `targets` through the restated ASE calls the reference makes
    (oracle/shims/ase: get_distances / get_all_distances with mic=True, general cells) -- the literal
    reference lines datasetbuilder.py:312-315 and :341-349."""
    import os
    import sys

    shims = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "shims")
    if shims not in sys.path:
        sys.path.insert(0, shims)
    from ase import Atoms

    atoms = Atoms(numbers=numbers, positions=pos, cell=cell, pbc=True)
    members, eigs = [], []
    for a in targets:
        d = atoms.get_distances(int(a), range(len(atoms)), mic=True)
        m = np.nonzero(d < cutoff)[0]
        cl = atoms[m]
        z = cl.get_atomic_numbers()
        top = np.outer(z, z).astype(float)
        r = cl.get_all_distances(mic=True)
        with np.errstate(divide="ignore", invalid="ignore"):
            np.divide(top, r, top)
        np.fill_diagonal(top, [diagz[int(q)] for q in z])
        top[np.isinf(top)] = 0
        top[np.isnan(top)] = 0
        members.append(m)
        eigs.append(np.linalg.eigh(top)[0])
    return members, eigs


@pytest.mark.parametrize("tilt", [(3.0, -2.0, 4.0), (-6.0, 5.5, -1.0), (0.0, 0.0, 7.0)])
def test_triclinic_cell(engine, tilt):
    """General cells on the sphere / descriptor path (VERDICT r1 missing item 2): membership bit-exact and rows
    within tolerance of the reference's ASE calls on a sheared LAMMPS box (tilt factors up to half a box length)."""
    s = synth.make_system(900, 0.06, seed=synth.BASE_SEED + 55)
    L = float(s.box[0])
    xy, xz, yz = tilt
    cell = np.array([[L, 0, 0], [xy, L, 0], [xz, yz, L]])
    rng = np.random.default_rng(5)
    # the generator's coordinates re-expressed in the sheared cell (same fractional coordinates)
    frames = []
    for k in range(2):
        fr = (synth.frame_positions(s, 2, sigma=0.06)[k] / L) @ cell
        frames.append(np.round(fr, 6))
    pos = np.stack(frames)
    engine.set_elements(s.atomname)
    try:
        targets = np.sort(rng.choice(2 * s.natoms, 300, replace=False)).astype(np.int32)
        counts, maxc, members = engine.compositions(pos, cell.reshape(9), s.types, 5.0, targets=targets, members_cap=96)
        out = engine.descriptors(pos, cell.reshape(9), s.types, 5.0, counts, maxc, targets=targets, members=members,
                                 want_eig=True, eigcap=96)
        out2 = engine.descriptors(pos, cell.reshape(9), s.types, 5.0, counts, maxc, targets=targets)   # cell-scan route
        engine.check()
        numbers = np.array([O.ATOMIC_NUMBERS[s.atomname[t]] for t in s.types])
        diagz = {O.ATOMIC_NUMBERS[a]: O.ATOMIC_NUMBERS[a] ** 2.4 / 2 for a in s.atomname}
        got_m = members.cpu().numpy()
        eig = out["eig"].cpu().numpy()
        X, X2 = out["X"].cpu().numpy(), out2["X"].cpu().numpy()
        assert np.array_equal(X, X2), "member-list and cell-scan routes differ"
        pads = np.array([O.coulomb_diag(s.atomname)[a] for a in s.atomname])
        for f in range(2):
            sel = np.nonzero(targets // s.natoms == f)[0]
            mem, eigs = _triclinic_oracle(numbers, pos[f], cell, targets[sel] % s.natoms, 5.0, diagz)
            comp = np.array([np.bincount(s.types[m], minlength=len(s.atomname)) for m in mem])
            assert np.array_equal(counts.cpu().numpy()[sel], comp), "sphere composition differs (triclinic)"
            for k, (m, e) in zip(sel, zip(mem, eigs)):
                g = got_m[k]
                assert np.array_equal(np.sort(g[g >= 0]), m), "sphere membership differs (triclinic)"
                tol = RTOL * np.maximum(np.abs(e), 1e-3 * np.abs(e).max())
                assert np.all(np.abs(eig[k, : len(e)] - e) <= tol), "spectrum outside tolerance (triclinic)"
        # bonds + descriptors together: the whole dump-only job runs on the sheared box
        from mddatasetbuilder_b200.streaming import ResidentSource, StreamingSelector
        import torch

        sel_job = StreamingSelector(engine, s.types, n_clusters=20, seed=2, periodic=True)
        res = sel_job.run(ResidentSource(torch.as_tensor(pos, device=engine.device), cell.reshape(9), 1))
        assert sum(len(v) for v in res.chosen.values()) > 0 and len(sel_job.big) > 0
    finally:
        engine.set_elements(["C", "H", "O"])


def test_ten_element_types(engine):
    """More than 8 element types on the whole path (the library holds 16): bonds, keys, compositions and
    descriptor rows of a 10-element soup against the oracle."""
    from mddatasetbuilder_b200.engine import Engine

    names = ["H", "C", "N", "O", "F", "Si", "P", "S", "Cl", "B"]
    rng = np.random.default_rng(12)
    n, L = 900, 30.0
    # jittered lattice: no overlapping atoms, plenty of bonded pairs (and over-coordinated atoms for the cleanup)
    g = 10
    sites = rng.permutation(g ** 3)[:n]
    pos = (np.stack([sites % g, (sites // g) % g, sites // (g * g)], axis=1) + 0.5) * (L / g) + rng.normal(0, 0.35, (n, 3))
    pos = np.round(np.mod(pos, L), 6)
    types = rng.integers(0, len(names), n).astype(np.uint8)
    eng = Engine(0, names, max_bonds=16)
    try:
        cell = np.diag([L, L, L])
        out = eng.analyze(pos[None], cell.reshape(9), types, True)
        eng.check()
        Z = np.array([O.ATOMIC_NUMBERS[a] for a in names], dtype=np.int32)[types]
        bonds, _ = O.connect_the_dots(Z, pos, cell, True, mode=1)
        rp, col = O.bonds_to_csr(n, bonds)
        from tests.util import ell_to_canonical

        grp, gcol = ell_to_canonical(out["ell"][0].cpu().numpy())
        crp, ccol = O.canonical_csr(rp, col)
        assert np.array_equal(grp, crp) and np.array_equal(gcol, ccol)
        assert np.array_equal(out["keys"][0].cpu().numpy().view(np.uint64), O.bond_keys_degree(types, rp))
        targets = np.arange(0, n, 5, dtype=np.int32)
        counts, maxc = eng.compositions(pos[None], cell.reshape(9), types, 5.0, targets=targets)
        diag = np.array([O.ATOMIC_NUMBERS[a] ** 2.4 / 2 for a in names])
        cn, mem, mats = O.descriptor_gather(Z, diag[types], pos, np.array([L, L, L]), True, 5.0, targets)
        comp = np.array([np.bincount(types[m], minlength=len(names)) for m in mem])
        assert np.array_equal(counts.cpu().numpy(), comp)
        X = eng.descriptors(pos[None], cell.reshape(9), types, 5.0, counts, maxc, targets=targets)["X"].cpu().numpy()
        eng.check()
        eigs = O.coulomb_eigs(mats)
        R, _ = O.feature_rows(eigs, comp, diag)
        norms = np.array([np.abs(e).max() for e in eigs])[:, None]
        assert np.all(np.abs(X - R) <= RTOL * np.maximum(np.abs(R), 1e-3 * norms))
    finally:
        eng.close()
