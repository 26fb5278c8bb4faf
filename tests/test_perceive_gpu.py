"""GPU parity: bond-order perception (csrc/perceive.cu) against the restatement in oracle/perceive.py,
slot for slot, on synthetic periodic frames and on hand-built molecules (reference detect.py:279-280;
Open Babel itself is absent from this image -- PARITY UNPINNED beyond the restatement)."""
import math
import os
import sys

import numpy as np
import pytest

from mddatasetbuilder_b200 import synth
from oracle.perceive import perceive_bond_orders

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
SHIMS = os.path.join(os.path.dirname(HERE), "oracle", "shims")
ZNUM = {"C": 6, "H": 1, "O": 8, "N": 7}


def _uc(cell):
    if SHIMS not in sys.path:
        sys.path.insert(0, SHIMS)
    from openbabel import openbabel

    uc = openbabel.OBUnitCell()
    m = np.asarray(cell, dtype=float).reshape(3, 3)
    uc.SetData(*(openbabel.vector3(*m[k]) for k in range(3)))
    return uc


def _oracle_orders(ell, P, Z, cell, periodic):
    """orders per ELL slot from the restatement, for the bonds the device found."""
    N = len(Z)
    nbrs = []
    for i in range(N):
        nb = [int(j) for j in ell[i] if j >= 0]
        nb.sort(key=lambda j: (P[j, 2], j))  # Open Babel's per-atom bond order
        nbrs.append(nb)
    uc = _uc(cell) if periodic else None

    def delta(a, b):
        d = (P[a] - P[b]).reshape(1, 3)
        if uc is not None:
            d = uc.mic(d)
        return (float(d[0, 0]), float(d[0, 1]), float(d[0, 2]))

    orders, info = perceive_bond_orders(Z, nbrs, delta)
    out = np.zeros(ell.shape, dtype=np.uint8)
    for i in range(N):
        for k, j in enumerate(ell[i]):
            if j >= 0:
                out[i, k] = orders[i][nbrs[i].index(int(j))]
    return out, info


def _key(tidx, row):
    lv = sorted(int(o) for o in row if o)
    key = ((tidx + 1) << 56) | (len(lv) << 48)
    for x, v in enumerate(lv[:12]):
        key |= v << (4 * x)
    return key


@pytest.mark.parametrize("natoms,density,sigma,seed", [(1500, 0.03, 0.03, 7), (2500, 0.06, 0.06, 8), (900, 0.10, 0.10, 9)])
def test_synthetic_frames_match_the_restatement(engine, natoms, density, sigma, seed):
    s = synth.make_system(natoms, density, seed=synth.BASE_SEED + 200 + seed, reactive_fraction=0.05)
    frames = synth.frame_positions(s, 3, sigma=sigma)
    cell = np.diag(s.box).reshape(9)
    types = s.types.astype(np.uint8)
    Z = [ZNUM[s.atomname[t]] for t in types]
    out = engine.analyze(frames, cell, types)
    engine.check()
    got = engine.perceive(frames, cell, types, out["ell"], labels=out["labels"])
    ring = engine.perceive_stats()["rings"]
    engine.check()
    ell = out["ell"].cpu().numpy()
    orders = got["orders"].cpu().numpy()
    keys = got["keys"].cpu().numpy().view(np.uint64)
    rings = 0
    nmult = 0
    for f in range(frames.shape[0]):
        want, info = _oracle_orders(ell[f], frames[f], Z, cell, True)
        rings += info["rings"]
        bad = np.argwhere(want != orders[f])
        assert bad.size == 0, f"frame {f}: {len(bad)} slots differ, first at atom {bad[0][0]}"
        nmult += int((want > 1).sum())
        wk = np.array([_key(int(types[i]), want[i]) for i in range(len(Z))], dtype=np.uint64)
        assert np.array_equal(wk, keys[f])
    assert ring == rings
    assert nmult > 0, "the sample must exercise multiple bonds"


def test_wide_rows_and_triclinic_cell(engine):
    """max_bonds = 16 rows (the remedy for crowded frames) and a tilted cell go through the same kernels."""
    s = synth.make_system(1800, 0.07, seed=synth.BASE_SEED + 244, reactive_fraction=0.1)
    rng = np.random.default_rng(3)
    pos = s.pos0 + rng.normal(0, 0.08, s.pos0.shape)
    L = s.box[0]
    cell = np.array([[L, 0, 0], [0.3 * L, L, 0], [-0.2 * L, 0.25 * L, L]])
    p = (pos / L) @ cell
    types = s.types.astype(np.uint8)
    Z = [ZNUM[s.atomname[t]] for t in types]
    engine.set_option("max_bonds", 16)
    try:
        out = engine.analyze(p[None], cell, types)
        engine.check()
        got = engine.perceive(p[None], cell, types, out["ell"], labels=out["labels"])
        rings = engine.perceive_stats()["rings"]
        engine.check()
    finally:
        engine.set_option("max_bonds", 8)
    ell = out["ell"].cpu().numpy()[0]
    assert ell.shape[1] == 16
    want, info = _oracle_orders(ell, p, Z, cell.reshape(-1), True)
    assert np.array_equal(want, got["orders"].cpu().numpy()[0])
    assert rings == info["rings"] and int((want > 1).sum()) > 0


def _run_molecule(engine, spec):
    names = ["C", "H", "O"]
    P = np.array([s[1:] for s in spec], dtype=float) + 20.0
    types = np.array([names.index(s[0]) for s in spec], dtype=np.uint8)
    out = engine.analyze(P[None], np.diag([60.0, 60.0, 60.0]).reshape(9), types, periodic=False)
    engine.check()
    got = engine.perceive(P[None], None, types, out["ell"], periodic=False, labels=out["labels"])
    engine.check()
    ell = out["ell"].cpu().numpy()[0]
    want, _ = _oracle_orders(ell, P, [ZNUM[s[0]] for s in spec], None, False)
    orders = got["orders"].cpu().numpy()[0]
    assert np.array_equal(want, orders)
    return [sorted(int(o) for o in row if o) for row in orders]


def test_atoms_without_bonds(engine):
    """Isolated atoms (the reference's key "O", "H": element with no levels) and the O2 known answers."""
    assert _run_molecule(engine, [("O", 0, 0, 0)]) == [[]]
    assert _run_molecule(engine, [("H", 0, 0, 0), ("C", 5.0, 0, 0), ("O", 0, 7.0, 0)]) == [[], [], []]
    assert _run_molecule(engine, [("O", 0, 0, 0), ("O", 1.0, 1.0, 1.0)]) == [[1], [1]]  # reference tests/test_bond.py:19-20
    assert _run_molecule(engine, [("O", 0, 0, 0), ("O", 1.21, 0, 0)]) == [[2], [2]]


def test_hand_built_molecules(engine):
    s, c = 1.087 * math.sin(math.radians(121.3)), 1.087 * math.cos(math.radians(121.3))
    lv = _run_molecule(engine, [("C", 0, 0, 0), ("C", 1.339, 0, 0), ("H", c, s, 0), ("H", c, -s, 0), ("H", 1.339 - c, s, 0),
                                ("H", 1.339 - c, -s, 0)])
    assert lv[0] == [1, 1, 2] and lv[1] == [1, 1, 2]
    lv = _run_molecule(engine, [("H", -1.06, 0, 0), ("C", 0, 0, 0), ("C", 1.203, 0, 0), ("H", 2.263, 0, 0)])
    assert lv[1] == [1, 3] and lv[2] == [1, 3]
    lv = _run_molecule(engine, [("O", -1.16, 0, 0), ("C", 0, 0, 0), ("O", 1.16, 0, 0)])
    assert lv == [[2], [2, 2], [2]]
    lv = _run_molecule(engine, [("C", 0, 0, 0), ("O", 1.128, 0, 0)])
    assert lv == [[2], [2]]
    s2, c2 = 1.11 * math.sin(math.radians(121.9)), 1.11 * math.cos(math.radians(121.9))
    lv = _run_molecule(engine, [("C", 0, 0, 0), ("O", 1.205, 0, 0), ("H", c2, s2, 0), ("H", c2, -s2, 0)])
    assert lv[0] == [1, 1, 2] and lv[1] == [2]
    # formic acid
    spec = [("C", 0, 0, 0), ("O", 1.20 * math.cos(math.radians(30)), 1.20 * math.sin(math.radians(30)), 0),
            ("O", 1.34 * math.cos(math.radians(150)), 1.34 * math.sin(math.radians(150)), 0), ("H", 0, -1.10, 0),
            ("H", 1.34 * math.cos(math.radians(150)) - 0.3, 1.34 * math.sin(math.radians(150)) + 0.92, 0)]
    lv = _run_molecule(engine, spec)
    assert lv[0] == [1, 1, 2] and lv[1] == [2] and lv[2] == [1, 1]
    # allene
    s3, c3 = 1.087 * math.sin(math.radians(120.9)), 1.087 * math.cos(math.radians(120.9))
    lv = _run_molecule(engine, [("C", -1.308, 0, 0), ("C", 0, 0, 0), ("C", 1.308, 0, 0), ("H", -1.308 + c3, s3, 0),
                                ("H", -1.308 + c3, -s3, 0), ("H", 1.308 - c3, 0, s3), ("H", 1.308 - c3, 0, -s3)])
    assert lv[1] == [2, 2] and lv[0] == [1, 1, 2]
    # s-trans butadiene
    spec = [("C", -1.85, 0.55, 0), ("C", -0.60, 0.0, 0), ("C", 0.60, 0.75, 0), ("C", 1.85, 0.20, 0), ("H", -2.75, -0.05, 0),
            ("H", -2.0, 1.62, 0), ("H", -0.50, -1.08, 0), ("H", 0.50, 1.83, 0), ("H", 2.75, 0.80, 0), ("H", 2.0, -0.87, 0)]
    lv = _run_molecule(engine, spec)
    assert [lv[k] for k in range(4)] == [[1, 1, 2]] * 4


def test_analyze_option_returns_perceived_keys(engine):
    s = synth.make_system(1200, 0.04, seed=synth.BASE_SEED + 231)
    frames = synth.frame_positions(s, 2, sigma=0.03)
    cell = np.diag(s.box).reshape(9)
    types = s.types.astype(np.uint8)
    plain = engine.analyze(frames, cell, types)
    want = engine.perceive(frames, cell, types, plain["ell"], labels=plain["labels"])["keys"]
    engine.set_option("perceive_bond_orders", 1)
    try:
        out = engine.analyze(frames, cell, types)
        engine.check()
    finally:
        engine.set_option("perceive_bond_orders", 0)
    assert np.array_equal(out["keys"].cpu().numpy(), want.cpu().numpy())
    assert np.array_equal(out["ell"].cpu().numpy(), plain["ell"].cpu().numpy())
    assert np.array_equal(out["labels"].cpu().numpy(), plain["labels"].cpu().numpy())
    assert not np.array_equal(out["keys"].cpu().numpy(), plain["keys"].cpu().numpy())
    # the host-buffer pipeline (chunks of frames on internal streams) honours the option as well
    engine.set_option("perceive_bond_orders", 1)
    engine.set_option("frames_per_launch", 1)
    try:
        host = engine.analyze_host(frames, cell, types)
        coded = engine.analyze_host(frames, cell, types, key_codes=True)
        engine.check()
    finally:
        engine.set_option("perceive_bond_orders", 0)
        engine.set_option("frames_per_launch", 0)
    assert np.array_equal(np.asarray(host["keys"]).view(np.int64), want.cpu().numpy())
    table = np.asarray(coded["keytable"]).view(np.uint64)
    assert np.array_equal(table[np.asarray(coded["keycode"])].view(np.int64), want.cpu().numpy())
