"""CPU restatement of Open Babel 3.1.x ``OBMol::PerceiveBondOrders`` for the C/H/O(/N) subset.

TEST INFRASTRUCTURE ONLY (imported by tests/, the openbabel stand-in under oracle/shims and the
golden generator) -- never by the product path.

The reference calls it at detect.py:279-280 (dump-only mode, ``readlevel=True``); its source lives in
the third-party dependency ``openbabel-wheel >= 3.1.0.0`` (reference pyproject.toml:38), which is in
neither /root/reference nor this image.  What follows restates the published algorithm
(openbabel ``src/mol.cpp`` OBMol::PerceiveBondOrders, ``src/bondtyper.cpp``
OBBondTyper::AssignFunctionalGroupBonds, ``data/bondtyp.txt``, ``src/atom.cpp`` AverageBondAngle,
``src/bond.cpp`` IsDoubleBondGeometry, ``src/math/vector3.cpp`` CalcTorsionAngle / vectorAngle)
FROM MEMORY.  PARITY UNPINNED: the only known answer the reference holds is O2 at 1.732 A -> order
1 (tests/test_bond.py:19-20); the hand-derived cases in tests/test_perceive_cpu.py pin the
restatement against itself, not against Open Babel.

Scope (stated deviations, the same on the device, csrc/perceive.cu):
  * pass 2 (planar 5/6-rings -> sp2) and pass 5 (aromatic candidates -> Kekule) are NOT restated:
    molecules that contain a cycle go through passes 1, 3, 4, 6 only (``IsInRing`` treated as
    false); ``rings`` reports how many independent cycles the frame holds;
  * pass 4 keeps the bondtyp.txt patterns that can match C/H/O molecules and are needed because
    pass 6 cannot reach them -- carboxylic acid / ester, carbon dioxide, allene -- and the coded
    carbonyl rule of AssignFunctionalGroupBonds; patterns that need N, P, S, Se or metals, and the
    "ene-one" pattern (redundant with carbonyl + pass 6 for sane geometries), are not restated;
  * std::sort in pass 6 is not stable; equal keys (the two atoms of one homonuclear bond) are
    ordered by atom index here.
"""
import math

RCOV = [0.0, 0.31, 0.28, 1.28, 0.96, 0.84, 0.76, 0.71, 0.66, 0.57, 0.58, 1.66, 1.41, 1.21, 1.11, 1.07,
        1.05, 1.02, 1.06]
MAXBONDS = [0, 1, 0, 1, 2, 4, 4, 4, 2, 1, 0, 1, 2, 6, 6, 6, 6, 1, 0]
# Pauling electronegativities, OBElements::GetElectroNeg
ELNEG = [0.0, 2.20, 0.0, 0.98, 1.57, 2.04, 2.55, 3.04, 3.44, 3.98, 0.0, 0.93, 1.31, 1.61, 1.90, 2.19,
         2.58, 3.16, 0.0]
RAD_TO_DEG = 180.0 / math.pi


def _dot(a, b):
    return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]


def _cross(a, b):
    return (a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0])


def vector_angle(v1, v2):
    """vector3.cpp vectorAngle (degrees)."""
    dp = _dot(v1, v2) / (math.sqrt(_dot(v1, v1)) * math.sqrt(_dot(v2, v2)))
    if dp < -0.999999:
        dp = -0.9999999
    if dp > 0.9999999:
        dp = 0.9999999
    return RAD_TO_DEG * math.acos(dp)


def torsion_angle(b1, b2, b3):
    """vector3.cpp CalcTorsionAngle with b1 = a-b, b2 = b-c, b3 = c-d (degrees)."""
    rb2 = math.sqrt(_dot(b2, b2))
    b2xb3 = _cross(b2, b3)
    b1xb2 = _cross(b1, b2)
    s = tuple(rb2 * x for x in b1)
    return -math.atan2(_dot(s, b2xb3), _dot(b1xb2, b2xb3)) * RAD_TO_DEG


def perceive_bond_orders(Z, nbrs, delta):
    """Bond orders after PerceiveBondOrders.

    Z      atomic numbers, one per atom
    nbrs   per atom, the bonded partners in Open Babel's per-atom bond order (ascending partner
           z-rank, what ConnectTheDots leaves behind)
    delta  delta(a, b) -> minimum-image vector(a) - vector(b) as a 3-tuple of floats

    Returns (orders, info): orders[i][k] is the order of the bond from atom i to nbrs[i][k];
    info = {"hyb": [...], "rings": n}.
    """
    n = len(Z)
    order = [[1] * len(nb) for nb in nbrs]
    deg = [len(nb) for nb in nbrs]

    def slot(i, j):
        return nbrs[i].index(j)

    def get(i, j):
        return order[i][slot(i, j)]

    def put(i, j, o):
        order[i][slot(i, j)] = o
        order[j][slot(j, i)] = o

    def length(i, j):
        v = delta(i, j)
        return math.sqrt(_dot(v, v))

    def valence(i):
        return sum(order[i])

    def nonsingle(i):
        return any(o != 1 for o in order[i])

    def avg_angle(i):
        """atom.cpp OBAtom::AverageBondAngle (periodic-aware GetAngle, 3.1)."""
        tot, cnt = 0.0, 0
        for x in range(deg[i]):
            v1 = delta(nbrs[i][x], i)
            for y in range(x + 1, deg[i]):
                v2 = delta(nbrs[i][y], i)
                l1, l2 = math.sqrt(_dot(v1, v1)), math.sqrt(_dot(v2, v2))
                if abs(l1) < 1e-3 or abs(l2) < 1e-3:
                    d = 0.0
                else:
                    d = vector_angle(v1, v2)
                tot += d
                cnt += 1
        return tot / cnt if cnt >= 1 else 0.0

    # pass 1: hybridisation from the average bond angle (mol.cpp)
    ang = [avg_angle(i) for i in range(n)]
    hyb = [0] * n
    for i in range(n):
        if ang[i] > 155.0:
            hyb[i] = 1
        elif ang[i] > 115.0:
            hyb[i] = 2
        if Z[i] == 7 and deg[i] == 2 and sum(1 for j in nbrs[i] if Z[j] == 1) == 1 and ang[i] > 109.5:
            hyb[i] = 2  # imine special case

    # pass 3: "antialiasing" (a no-op unless a neighbour already holds hyb 3, restated literally)
    for i in range(n):
        if hyb[i] in (1, 2):
            open_nbr = any(hyb[j] < 3 or deg[j] == 1 for j in nbrs[i])
            if not open_nbr:
                hyb[i] = 3 if hyb[i] == 2 else 2

    # pass 4: functional groups.  Every pattern is matched on the state left by the previous one
    # (implicit SMARTS bonds only match single bonds), all matches first, then the assignments.
    # -- carboxylic acid, ester:  [#6D3^2]([#8D1])([#8])*   0 1 2  0 2 1  0 3 1
    todo = []
    for c in range(n):
        if Z[c] == 6 and deg[c] == 3 and hyb[c] == 2 and all(o == 1 for o in order[c]):
            for o1 in nbrs[c]:
                if Z[o1] == 8 and deg[o1] == 1:
                    rest = [x for x in nbrs[c] if x != o1]
                    o2 = next((x for x in rest if Z[x] == 8), None)
                    if o2 is not None:
                        r = next(x for x in rest if x != o2)
                        todo.append((c, o1, o2, r))
                        break  # further mappings cover the same atom set (GetUMapList)
    for c, o1, o2, r in todo:
        put(c, o1, 2)
        put(c, o2, 1)
        put(c, r, 1)
    # -- carbon dioxide:  [#8D1][#6D2^1][#8D1]   0 1 2  1 2 2
    todo = []
    for c in range(n):
        if Z[c] == 6 and deg[c] == 2 and hyb[c] == 1 and all(o == 1 for o in order[c]):
            a, b = nbrs[c]
            if Z[a] == 8 and deg[a] == 1 and Z[b] == 8 and deg[b] == 1:
                todo.append((c, a, b))
    for c, a, b in todo:
        put(c, a, 2)
        put(c, b, 2)
    # -- allene:  [#6^2][#6D2^1][#6^2]   0 1 2  1 2 2
    todo = []
    for c in range(n):
        if Z[c] == 6 and deg[c] == 2 and hyb[c] == 1 and all(o == 1 for o in order[c]):
            a, b = nbrs[c]
            if Z[a] == 6 and hyb[a] == 2 and Z[b] == 6 and hyb[b] == 2:
                todo.append((c, a, b))
    for c, a, b in todo:
        put(c, a, 2)
        put(c, b, 2)
    # -- carbonyl (bondtyper.cpp, coded):  [#8D1;!-][#6](*)(*), 115 < angle(C) < 150, d < 1.28
    todo = []
    for o in range(n):
        if Z[o] == 8 and deg[o] == 1:
            c = nbrs[o][0]
            if Z[c] == 6 and get(o, c) == 1:
                others = [x for x in nbrs[c] if x != o and get(c, x) == 1]
                if len(others) >= 2:
                    todo.append((o, c))
    for o, c in todo:
        a = avg_angle(c)
        if 115 < a < 150 and length(o, c) < 1.28:
            if not any(x == 2 for x in order[o]):
                put(o, c, 2)

    # pass 6: remaining multiple bonds, atoms by electronegativity, then shortest heavy-atom bond
    def corrected_rad(z, h):
        r = RCOV[z]
        return r * 0.95 if h == 2 else (r * 0.90 if h == 1 else r)

    def approx(a, b, prec):
        return abs(a - b) <= prec * min(abs(a), abs(b))

    def double_bond_geometry(a, b):
        """bond.cpp OBBond::IsDoubleBondGeometry."""
        if hyb[a] == 1 or deg[a] > 3 or hyb[b] == 1 or deg[b] > 3:
            return True
        b2 = delta(a, b)
        for s in nbrs[a]:
            if s == b:
                continue
            b1 = delta(s, a)
            for e in nbrs[b]:
                if e == a:
                    continue
                b3 = delta(b, e)
                t = abs(torsion_angle(b1, b2, b3))
                if 15.0 < t < 160.0:
                    return False
        return True

    keys = []
    for i in range(n):
        shortest = 1.0e5
        for j in nbrs[i]:
            if Z[j] != 1:
                shortest = min(shortest, length(i, j))
        keys.append((ELNEG[Z[i]] * 1e6 + shortest, i))
    keys.sort()
    for _, a in keys:
        mx = MAXBONDS[Z[a]]
        if (hyb[a] == 1 or deg[a] == 1) and valence(a) + 2 <= mx:
            if nonsingle(a) or (Z[a] == 7 and valence(a) + 2 > 3):
                continue
            max_en, shortest, c = 0.0, 5000.0, None
            for b in nbrs[a]:
                en = ELNEG[Z[b]]
                if ((hyb[b] == 1 or deg[b] == 1) and valence(b) + 2 <= MAXBONDS[Z[b]]
                        and (en > max_en or (approx(en, max_en, 1.0e-6) and length(a, b) < shortest))):
                    if nonsingle(b) or (Z[b] == 7 and valence(b) + 2 > 3):
                        continue
                    bl = length(a, b)
                    if deg[a] == 1 or deg[b] == 1:
                        if bl > 0.9 * (corrected_rad(Z[a], hyb[a]) + corrected_rad(Z[b], hyb[b])):
                            continue
                    shortest, max_en, c = bl, en, b
            if c is not None:
                put(a, c, 3)
        elif (hyb[a] == 2 or deg[a] == 1) and valence(a) + 1 <= mx:
            if nonsingle(a) or (Z[a] == 7 and valence(a) + 1 > 3):
                continue
            max_en, shortest, c = 0.0, 5000.0, None
            for b in nbrs[a]:
                en = ELNEG[Z[b]]
                if ((hyb[b] == 2 or deg[b] == 1) and valence(b) + 1 <= MAXBONDS[Z[b]]
                        and double_bond_geometry(a, b) and (en > max_en or approx(en, max_en, 1.0e-6))):
                    if nonsingle(b) or (Z[b] == 7 and valence(b) + 1 > 3):
                        continue
                    bl = length(a, b)
                    if deg[a] == 1 or deg[b] == 1:
                        if bl > 0.93 * (corrected_rad(Z[a], hyb[a]) + corrected_rad(Z[b], hyb[b])):
                            continue
                    diff = shortest - bl
                    # no ring perception here: IsInRing() is false, which leaves `diff > -0.01`
                    if diff > 0.1 or diff > -0.01:
                        shortest, max_en, c = bl, en, b
            if c is not None:
                put(a, c, 2)

    # molecules with a cycle (ring passes not restated): bonds >= atoms within the component
    parent = list(range(n))

    def root(x):
        while parent[x] != x:
            parent[x] = parent[parent[x]]
            x = parent[x]
        return x

    for i in range(n):
        for j in nbrs[i]:
            ri, rj = root(i), root(j)
            if ri != rj:
                parent[max(ri, rj)] = min(ri, rj)
    tally = {}
    for i in range(n):
        r = root(i)
        tally[r] = tally.get(r, 0) + deg[i] - 2
    # independent cycles = bonds - atoms + molecules (what the device sums, csrc/perceive.cu)
    rings = sum(v + 2 for v in tally.values()) // 2
    return order, {"hyb": hyb, "rings": rings}
