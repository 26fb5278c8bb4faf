"""GPU parity: bond detection / keys / connectivity vs the CPU oracle (bit-exact).

Mirrors the reference's own tests/test_bond.py (O2 known answers) and extends them with
seeded synthetic frames, through the C ABI (ctypes -> libmdb_b200.so)."""
import numpy as np
import pytest

from mddatasetbuilder_b200 import synth
from oracle import oracle as O
from tests.util import assert_ell_rows_sorted_padded, ell_to_canonical, oracle_frame

pytestmark = pytest.mark.gpu


def _run(engine, pos, cell, types, periodic=True):
    out = engine.analyze(pos, cell, types, periodic)
    stats = engine.check()
    return out, stats


def test_o2_pbc(engine):
    # reference tests/test_bond.py:9-20
    engine.set_elements(["O"])
    pos = np.array([[[0.0, 0.0, 0.0], [19.0, 19.0, 19.0]]])
    out, _ = _run(engine, pos, np.diag([20.0, 20.0, 20.0]), np.zeros(2, dtype=np.uint8), True)
    ell = out["ell"].cpu().numpy()[0]
    assert ell[0, 0] == 1 and ell[1, 0] == 0 and (ell[:, 1:] == -1).all()
    assert out["labels"].cpu().numpy()[0].tolist() == [0, 0]
    keys = out["keys"].cpu().numpy().view(np.uint64)[0]
    assert [O.key_string(k, ["O"]) for k in keys] == ["O1", "O1"]
    engine.set_elements(["C", "H", "O"])


def test_o2_nopbc(engine):
    # reference tests/test_bond.py:23-34
    engine.set_elements(["O"])
    pos = np.array([[[0.0, 0.0, 0.0], [19.0, 19.0, 19.0]]])
    out, _ = _run(engine, pos, np.diag([20.0, 20.0, 20.0]), np.zeros(2, dtype=np.uint8), False)
    ell = out["ell"].cpu().numpy()[0]
    assert (ell == -1).all()
    assert out["labels"].cpu().numpy()[0].tolist() == [0, 1]
    keys = out["keys"].cpu().numpy().view(np.uint64)[0]
    assert [O.key_string(k, ["O"]) for k in keys] == ["O", "O"]
    engine.set_elements(["C", "H", "O"])


@pytest.mark.parametrize("natoms,density,dense", [(3000, 0.02, False), (3000, 0.05, False), (4000, 0.11, True),
                                                  (20000, 0.05, False)])
def test_synthetic_frames(engine, natoms, density, dense):
    s = synth.make_system(natoms, density, seed=synth.BASE_SEED + natoms, dense=dense)
    frames = synth.frame_positions(s, 4)
    out, stats = _run(engine, frames, s.cell(), s.types)
    ell = out["ell"].cpu().numpy()
    labels = out["labels"].cpu().numpy()
    keys = out["keys"].cpu().numpy().view(np.uint64)
    ndel_total = 0
    for f in range(len(frames)):
        rp, col, ndel = oracle_frame(s, frames[f], mode=1 if natoms > 5000 else 0)
        ndel_total += ndel
        assert_ell_rows_sorted_padded(ell[f])
        grp, gcol = ell_to_canonical(ell[f])
        crp, ccol = O.canonical_csr(rp, col)
        assert np.array_equal(grp, crp), f"degree mismatch in frame {f}"
        assert np.array_equal(gcol, ccol), f"bond list mismatch in frame {f}"
        _, _, lab = O.dps(rp, col)
        assert np.array_equal(labels[f], lab), f"molecule labels mismatch in frame {f}"
        assert np.array_equal(keys[f], O.bond_keys_degree(s.types, rp)), f"keys mismatch in frame {f}"
    assert ndel_total > 0, "the synthetic frames must exercise the valence/angle cleanup"
    assert stats["overflow"] == 0


def test_ob_order_and_dps(engine):
    s = synth.make_system(3000, 0.05, seed=synth.BASE_SEED + 5)
    frames = synth.frame_positions(s, 2)
    out, _ = _run(engine, frames, s.cell(), s.types)
    rowptr, col = engine.ell_to_csr(out["ell"], order=1, pos=out["pos"])
    rowptr = rowptr.cpu().numpy()
    col = col.cpu().numpy()
    for f in range(2):
        rp, c, _ = oracle_frame(s, frames[f], mode=0)
        assert np.array_equal(rowptr[f], rp)
        assert np.array_equal(col[f][: rp[-1]], c), "adjacency order differs from the Open Babel bond order"
        import torch

        ma, mp = engine.dps(torch.as_tensor(rp, device=engine.device), torch.as_tensor(c, device=engine.device))
        oma, omp, _ = O.dps(rp, c)
        assert np.array_equal(mp.cpu().numpy(), omp)
        assert np.array_equal(ma.cpu().numpy(), oma), "dps visiting order differs"
    # canonical order conversion
    rowptr0, col0 = engine.ell_to_csr(out["ell"], order=0)
    rp, c, _ = oracle_frame(s, frames[0], mode=0)
    crp, ccol = O.canonical_csr(rp, c)
    assert np.array_equal(rowptr0.cpu().numpy()[0], crp)
    assert np.array_equal(col0.cpu().numpy()[0][: crp[-1]], ccol)


def test_triclinic_and_small_box(engine):
    rng = np.random.default_rng(7)
    # LAMMPS-style lower-triangular tilted cell
    s = synth.make_system(2500, 0.05, seed=synth.BASE_SEED + 11)
    L = s.box[0]
    cell = np.array([[L, 0, 0], [0.3 * L, L, 0], [-0.2 * L, 0.25 * L, L]])
    # keep fractional coordinates, re-embed in the tilted cell
    fr = s.pos0 / L
    pos = fr @ cell
    out, _ = _run(engine, pos[None], cell, s.types)
    rp, col, _ = oracle_frame(s, pos, mode=0, cell=cell)
    grp, gcol = ell_to_canonical(out["ell"].cpu().numpy()[0])
    crp, ccol = O.canonical_csr(rp, col)
    assert np.array_equal(grp, crp) and np.array_equal(gcol, ccol)
    # box so small that every axis falls back to the single-cell min-image path
    s2 = synth.make_system(12, 0.08, seed=synth.BASE_SEED + 12)
    assert s2.box[0] < 3 * 2.0
    pos2 = s2.pos0 + rng.normal(0, 0.02, s2.pos0.shape)
    out2, _ = _run(engine, pos2[None], s2.cell(), s2.types)
    rp, col, _ = oracle_frame(s2, pos2, mode=0)
    grp, gcol = ell_to_canonical(out2["ell"].cpu().numpy()[0])
    crp, ccol = O.canonical_csr(rp, col)
    assert np.array_equal(grp, crp) and np.array_equal(gcol, ccol)


def test_nonperiodic(engine):
    s = synth.make_system(3000, 0.05, seed=synth.BASE_SEED + 13)
    out, _ = _run(engine, s.pos0[None], s.cell(), s.types, periodic=False)
    rp, col, _ = oracle_frame(s, s.pos0, periodic=False, mode=0)
    grp, gcol = ell_to_canonical(out["ell"].cpu().numpy()[0])
    crp, ccol = O.canonical_csr(rp, col)
    assert np.array_equal(grp, crp) and np.array_equal(gcol, ccol)
    _, _, lab = O.dps(rp, col)
    assert np.array_equal(out["labels"].cpu().numpy()[0], lab)


def test_host_path_matches_device_path(engine):
    s = synth.make_system(5000, 0.05, seed=synth.BASE_SEED + 14)
    frames = synth.frame_positions(s, 7)
    engine.set_option("frames_per_launch", 2)  # force several pipeline chunks
    try:
        dev = engine.analyze(frames, s.cell(), s.types)
        engine.check()
        host = engine.analyze_host(frames, s.cell(), s.types)
    finally:
        engine.set_option("frames_per_launch", 0)
    assert np.array_equal(dev["ell"].cpu().numpy(), host["ell"])
    assert np.array_equal(dev["keys"].cpu().numpy().view(np.uint64), host["keys"])
    assert np.array_equal(dev["labels"].cpu().numpy(), host["labels"])
    # 4-slot rows (C/H/O: degree <= max valence 4 after the cleanup) carry the same bonds
    narrow = engine.analyze_host(frames, s.cell(), s.types, ell_width=4)
    full = dev["ell"].cpu().numpy()
    assert narrow["ell"].shape[-1] == 4 and np.array_equal(narrow["ell"], full[:, :, :4]) and (full[:, :, 4:] < 0).all()
    # one-byte key codes + dictionary decode to the same 64-bit keys
    engine.set_option("frames_per_launch", 3)
    try:
        coded = engine.analyze_host(frames, s.cell(), s.types, ell_width=4, key_codes=True)
    finally:
        engine.set_option("frames_per_launch", 0)
    table = coded["keytable"]
    assert np.array_equal(table[coded["keycode"]], host["keys"])
    used = table != np.uint64(0xFFFFFFFFFFFFFFFF)
    assert set(table[used].tolist()) == set(np.unique(host["keys"]).tolist())
    assert np.array_equal(coded["labels"], host["labels"]) and np.array_equal(coded["ell"], narrow["ell"])
    # the bond list (every bond once, i < j, by ascending i then j) holds exactly the bonds of the rows
    for per in (1, 3, 100):
        engine.set_option("frames_per_launch", per)
        try:
            el = engine.analyze_host(frames, s.cell(), s.types, edges=True)
        finally:
            engine.set_option("frames_per_launch", 0)
        ptr = el["edge_ptr"]
        assert ptr[0] == 0 and np.all(np.diff(ptr) >= 0)
        for f in range(frames.shape[0]):
            i, k = np.nonzero(full[f] >= 0)
            j = full[f][i, k]
            want = np.stack([i[j > i], j[j > i]], axis=1)
            assert np.array_equal(el["edges"][ptr[f]:ptr[f + 1]], want)
        assert np.array_equal(el["labels"], host["labels"]) and np.array_equal(el["keytable"][el["keycode"]], host["keys"])
    with pytest.raises(RuntimeError, match="edge capacity"):
        engine.analyze_host(frames, s.cell(), s.types, edges=True, out={"edges": np.empty((10, 2), dtype=np.int32)})


def test_bondfile_keys(engine):
    # bond-file mode keys (reference detect.py:112-116) from the generator topology
    import torch

    s = synth.make_system(3000, 0.05, seed=synth.BASE_SEED + 15)
    N = s.natoms
    ell = np.full((1, N, 8), -1, dtype=np.int32)
    bo = np.zeros((1, N, 8))
    for i in range(N):
        a, b = s.rowptr[i], s.rowptr[i + 1]
        ell[0, i, : b - a] = s.col[a:b]
        bo[0, i, : b - a] = s.bo[a:b]
    # half-way cases of Python's banker's rounding
    bo[0, 0, 0] = 0.5
    bo[0, 1, 0] = 1.5
    bo[0, 2, 0] = 2.5
    keys = engine.bond_keys_bo(torch.as_tensor(ell, device=engine.device), torch.as_tensor(bo, device=engine.device),
                               engine.to_device_types(s.types))
    flat = np.concatenate([bo[0, i, : s.rowptr[i + 1] - s.rowptr[i]] for i in range(N)])
    want = O.bond_keys_bondfile(s.types, s.rowptr, flat)
    assert np.array_equal(keys.cpu().numpy().view(np.uint64)[0], want)
    for i in range(50):
        lv = sorted(max(1, round(float(x))) for x in bo[0, i, : s.rowptr[i + 1] - s.rowptr[i]])
        assert O.unpack_key(want[i]) == (int(s.types[i]), lv)
