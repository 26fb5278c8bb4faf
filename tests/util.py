"""Helpers shared by the parity tests."""
import numpy as np

from oracle import oracle as O


def ell_to_canonical(ell_frame):
    """ELL [N, maxb] (numpy) -> canonical (row-sorted) CSR."""
    ell_frame = np.asarray(ell_frame)
    deg = (ell_frame >= 0).sum(axis=1)
    rowptr = np.zeros(len(ell_frame) + 1, dtype=np.int32)
    np.cumsum(deg, out=rowptr[1:])
    col = ell_frame[ell_frame >= 0].astype(np.int32)  # row-major order keeps rows contiguous
    return O.canonical_csr(rowptr, col)


def oracle_frame(sys, pos, periodic=True, mode=1, cell=None):
    Z = np.array([O.ATOMIC_NUMBERS[sys.atomname[t]] for t in sys.types], dtype=np.int32)
    bonds, ndel = O.connect_the_dots(Z, pos, sys.cell() if cell is None else cell, periodic, mode)
    rowptr, col = O.bonds_to_csr(len(Z), bonds)
    return rowptr, col, ndel


def assert_ell_rows_sorted_padded(ell_frame):
    e = np.asarray(ell_frame)
    valid = e >= 0
    # padding only at the end of each row
    assert np.all(valid[:, :-1] | ~valid[:, 1:]), "holes inside ELL rows"
    a, b = e[:, :-1], e[:, 1:]
    ok = ~(valid[:, :-1] & valid[:, 1:]) | (a < b)
    assert np.all(ok), "ELL rows are not strictly ascending"
    assert np.all(e[~valid] == -1)


def c1_inputs(outdir, natoms=300, nframes=200, seed=303):
    """The C1 trajectory of SURVEY.md 8(d) as LAMMPS text, regenerated deterministically from the seeded
    generator (tests/golden/c1_golden.json pins the sha256 of both files): ~300 C/H/O atoms, 200 frames,
    shuffled atom lines, per-frame bond orders; plus a DeePMD-style atomic model-deviation file (one header
    line, per frame 7 leading columns and one value per atom in the dump's LINE order).
    Returns (system, dump path, bonds path, model deviation path)."""
    import os

    from mddatasetbuilder_b200 import synth

    s = synth.make_system(natoms, 0.045, seed=synth.BASE_SEED + seed, reactive_fraction=0.05)
    frames = synth.frame_positions(s, nframes, sigma=0.05)
    dump = os.path.join(outdir, "dump.c1")
    bonds = os.path.join(outdir, "bonds.c1")
    err = os.path.join(outdir, "model_devi.c1")
    synth.write_lammps_dump(dump, s, frames, shuffle_seed=seed)
    rng = np.random.default_rng(seed)
    nb = len(s.order)
    bo_frames = []
    for _ in range(nframes):
        per_bond = np.round(np.clip(s.order + rng.uniform(-0.3, 0.3, nb), 0.05, None), 3)
        bo_frames.append(per_bond[s.eid])
    synth.write_lammps_bonds(bonds, s, nframes, bo_frames=bo_frames)
    # model deviation rows follow the dump's line order: the same permutations write_lammps_dump drew
    prng = np.random.Generator(np.random.PCG64(seed))
    erng = np.random.default_rng(seed + 1)
    with open(err, "w") as f:
        f.write("#       step         max_devi_v         min_devi_v         avg_devi_v         max_devi_f         min_devi_f         avg_devi_f  atm_devi_f(N)\n")
        for k in range(nframes):
            order = prng.permutation(s.natoms)
            by_id = erng.random(s.natoms)
            f.write(" ".join(["%d" % k] + ["%.6e" % v for v in erng.random(6)] + ["%.6e" % by_id[i] for i in order]) + "\n")
    return s, dump, bonds, err
