"""The PerceiveBondOrders restatement (oracle/perceive.py) on molecules whose answer follows from the
published algorithm by hand.  PARITY UNPINNED against Open Babel itself (absent from this image):
the one known answer the reference holds is O2 at 1.732 A -> single (tests/test_bond.py:19-20)."""
import math
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle", "shims"))

from oracle.perceive import perceive_bond_orders  # noqa: E402

Z = {"H": 1, "C": 6, "N": 7, "O": 8}


def mol(spec, bonds):
    """spec: [(symbol, x, y, z)], bonds: [(i, j)] -> (Z, nbrs in ascending partner z-rank, delta)."""
    P = np.array([s[1:] for s in spec], dtype=float)
    zs = [Z[s[0]] for s in spec]
    nbrs = [[] for _ in spec]
    for i, j in bonds:
        nbrs[i].append(j)
        nbrs[j].append(i)
    for i in range(len(spec)):
        nbrs[i].sort(key=lambda j: (P[j, 2], j))
    return zs, nbrs, (lambda a, b: tuple(P[a] - P[b]))


def levels(spec, bonds):
    zs, nbrs, delta = mol(spec, bonds)
    orders, info = perceive_bond_orders(zs, nbrs, delta)
    return [sorted(o) for o in orders], info


def test_o2_known_answer_of_the_reference():
    lv, _ = levels([("O", 0, 0, 0), ("O", 1.0, 1.0, 1.0)], [(0, 1)])  # 1.732 A
    assert lv == [[1], [1]]
    lv, _ = levels([("O", 0, 0, 0), ("O", 1.21, 0, 0)], [(0, 1)])
    assert lv == [[2], [2]]  # 1.21 <= 0.93 * (0.66 + 0.66)


def test_small_molecules():
    # water, hydroxyl, hydroperoxyl: nothing to assign
    lv, _ = levels([("O", 0, 0, 0), ("H", 0.96, 0, 0), ("H", -0.24, 0.93, 0)], [(0, 1), (0, 2)])
    assert lv == [[1, 1], [1], [1]]
    # carbon monoxide: the oxygen's turn makes it a double bond (terminal length test 0.93 * 1.42)
    lv, _ = levels([("C", 0, 0, 0), ("O", 1.128, 0, 0)], [(0, 1)])
    assert lv == [[2], [2]]
    # carbon dioxide: the bondtyp.txt pattern
    lv, info = levels([("O", -1.16, 0, 0), ("C", 0, 0, 0), ("O", 1.16, 0, 0)], [(0, 1), (1, 2)])
    assert lv == [[2], [2, 2], [2]] and info["hyb"][1] == 1
    # a bent O-C-O (130 degrees) is no CO2 and nothing else fires: C is sp2 but both O fail valence+1 after...
    a = math.radians(130 / 2)
    lv, _ = levels([("O", -1.2 * math.sin(a), 1.2 * math.cos(a), 0), ("C", 0, 0, 0), ("O", 1.2 * math.sin(a), 1.2 * math.cos(a), 0)],
                   [(0, 1), (1, 2)])
    assert lv[1] in ([1, 2], [2, 2])  # pass 6 decides; never a triple


def _ethylene():
    s, c = 1.087 * math.sin(math.radians(121.3)), 1.087 * math.cos(math.radians(121.3))
    spec = [("C", 0, 0, 0), ("C", 1.339, 0, 0), ("H", c, s, 0), ("H", c, -s, 0), ("H", 1.339 - c, s, 0), ("H", 1.339 - c, -s, 0)]
    return spec, [(0, 1), (0, 2), (0, 3), (1, 4), (1, 5)]


def test_hydrocarbons():
    lv, info = levels(*_ethylene())
    assert lv[0] == [1, 1, 2] and lv[1] == [1, 1, 2] and info["hyb"][:2] == [2, 2]
    # twisted by 90 degrees: IsDoubleBondGeometry refuses
    spec, bonds = _ethylene()
    spec[4] = ("H", spec[4][1], 0, spec[4][2])
    spec[5] = ("H", spec[5][1], 0, -spec[4][3])
    lv, _ = levels(spec, bonds)
    assert lv[0] == [1, 1, 1]
    # acetylene: triple
    lv, _ = levels([("H", -1.06, 0, 0), ("C", 0, 0, 0), ("C", 1.203, 0, 0), ("H", 2.263, 0, 0)], [(0, 1), (1, 2), (2, 3)])
    assert lv[1] == [1, 3] and lv[2] == [1, 3]
    # methane and the planar methyl radical: single bonds only
    t = 1.09 / math.sqrt(3)
    lv, _ = levels([("C", 0, 0, 0), ("H", t, t, t), ("H", -t, -t, t), ("H", -t, t, -t), ("H", t, -t, -t)],
                   [(0, 1), (0, 2), (0, 3), (0, 4)])
    assert lv[0] == [1, 1, 1, 1]
    lv, info = levels([("C", 0, 0, 0), ("H", 1.08, 0, 0), ("H", -0.54, 0.935, 0), ("H", -0.54, -0.935, 0)],
                      [(0, 1), (0, 2), (0, 3)])
    assert lv[0] == [1, 1, 1] and info["hyb"][0] == 2
    # allene: the centre is sp, both ends sp2 -> the bondtyp.txt pattern
    s = 1.087 * math.sin(math.radians(120.9))
    c = 1.087 * math.cos(math.radians(120.9))
    spec = [("C", -1.308, 0, 0), ("C", 0, 0, 0), ("C", 1.308, 0, 0), ("H", -1.308 + c, s, 0), ("H", -1.308 + c, -s, 0),
            ("H", 1.308 - c, 0, s), ("H", 1.308 - c, 0, -s)]
    lv, _ = levels(spec, [(0, 1), (1, 2), (0, 3), (0, 4), (2, 5), (2, 6)])
    assert lv[1] == [2, 2] and lv[0] == [1, 1, 2]


def test_carbonyls():
    # formaldehyde: coded carbonyl rule (angle 120, d = 1.205)
    s, c = 1.11 * math.sin(math.radians(121.9)), 1.11 * math.cos(math.radians(121.9))
    lv, _ = levels([("C", 0, 0, 0), ("O", 1.205, 0, 0), ("H", c, s, 0), ("H", c, -s, 0)], [(0, 1), (0, 2), (0, 3)])
    assert lv[0] == [1, 1, 2] and lv[1] == [2]
    # a C-O of 1.43 A on an sp2 carbon stays single in pass 4; pass 6 then tests 0.93 * (0.66 + 0.95 * 0.76)
    lv, _ = levels([("C", 0, 0, 0), ("O", 1.43, 0, 0), ("H", c, s, 0), ("H", c, -s, 0)], [(0, 1), (0, 2), (0, 3)])
    assert lv[1] == [1]
    # formic acid: carboxylic pattern -> C=O to the terminal oxygen, C-O(H) single
    spec = [("C", 0, 0, 0), ("O", 1.20 * math.cos(math.radians(30)), 1.20 * math.sin(math.radians(30)), 0),
            ("O", 1.34 * math.cos(math.radians(150)), 1.34 * math.sin(math.radians(150)), 0), ("H", 0, -1.10, 0),
            ("H", 1.34 * math.cos(math.radians(150)) - 0.3, 1.34 * math.sin(math.radians(150)) + 0.92, 0)]
    lv, _ = levels(spec, [(0, 1), (0, 2), (0, 3), (2, 4)])
    assert lv[0] == [1, 1, 2] and lv[1] == [2] and lv[2] == [1, 1]
    # formate (two terminal oxygens): the first one in the carbon's bond order gets the double bond
    spec = [("C", 0, 0, 0), ("O", 1.25 * math.cos(math.radians(25)), 1.25 * math.sin(math.radians(25)), 0.2),
            ("O", 1.25 * math.cos(math.radians(155)), 1.25 * math.sin(math.radians(155)), -0.2), ("H", 0, -1.10, 0)]
    lv, _ = levels(spec, [(0, 1), (0, 2), (0, 3)])
    assert lv[0] == [1, 1, 2] and lv[2] == [2] and lv[1] == [1]  # atom 2 has the lower z -> first in OB order


def test_butadiene_and_ring_count():
    # s-trans butadiene, planar: C1=C2-C3=C4
    pts = [(-1.85, 0.55, 0), (-0.60, 0.0, 0), (0.60, 0.75, 0), (1.85, 0.20, 0)]  # C-C 1.366 / 1.415 / 1.366
    spec = [("C",) + p for p in pts]
    spec += [("H", -2.75, -0.05, 0), ("H", -2.0, 1.62, 0), ("H", -0.50, -1.08, 0), ("H", 0.50, 1.83, 0),
             ("H", 2.75, 0.80, 0), ("H", 2.0, -0.87, 0)]
    bonds = [(0, 1), (1, 2), (2, 3), (0, 4), (0, 5), (1, 6), (2, 7), (3, 8), (3, 9)]
    lv, info = levels(spec, bonds)
    assert lv[0] == [1, 1, 2] and lv[1] == [1, 1, 2] and lv[2] == [1, 1, 2] and lv[3] == [1, 1, 2]
    assert info["rings"] == 0
    # cyclopropane skeleton counts as a ring molecule
    lv, info = levels([("C", 0, 0, 0), ("C", 1.5, 0, 0), ("C", 0.75, 1.3, 0)], [(0, 1), (1, 2), (0, 2)])
    assert info["rings"] == 1


def test_shim_uses_it():
    from openbabel import openbabel

    m = openbabel.OBMol()
    for k, (z, p) in enumerate([(8, (0, 0, 0)), (6, (1.16, 0, 0)), (8, (2.32, 0, 0))]):
        a = m.NewAtom(k)
        a.SetAtomicNum(z)
        a.SetVector(*p)
    m.ConnectTheDots()
    m.PerceiveBondOrders()
    assert sorted(b.GetBondOrder() for b in openbabel.OBMolBondIter(m)) == [2, 2]
