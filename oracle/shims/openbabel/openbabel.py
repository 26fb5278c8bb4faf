"""openbabel.openbabel stand-in: the API surface touched at reference detect.py:255-291.

Pure-Python/numpy restatement of Open Babel 3.1.x OBMol::ConnectTheDots,
written independently of oracle/oracle.c so the two can be cross-checked.  PerceiveBondOrders is the
C/H/O-subset restatement in oracle/perceive.py (ring passes and most bondtyp.txt patterns left out,
see its header); the reference's only known answer for it (O2 -> [[1],[1]], tests/test_bond.py:19-20)
holds.  Test infrastructure only -- PARITY UNPINNED beyond tests/test_bond.py."""
import math

import numpy as np

RCOV = [0.0, 0.31, 0.28, 1.28, 0.96, 0.84, 0.76, 0.71, 0.66, 0.57, 0.58, 1.66, 1.41, 1.21, 1.11, 1.07,
        1.05, 1.02, 1.06]
MAXBONDS = [0, 1, 0, 1, 2, 4, 4, 4, 2, 1, 0, 1, 2, 6, 6, 6, 6, 1, 0]


class vector3:
    def __init__(self, x=0.0, y=0.0, z=0.0):
        self.v = (float(x), float(y), float(z))


class OBUnitCell:
    def __init__(self):
        self.ortho = None

    def SetData(self, v1, v2, v3):
        m = np.array([v1.v, v2.v, v3.v], dtype=float)
        self.ortho = m.T.copy()  # _mOrtho: lattice vectors as columns
        e = self.ortho
        x = e[0][0] * (e[1][1] * e[2][2] - e[1][2] * e[2][1])
        y = e[0][1] * (e[1][2] * e[2][0] - e[1][0] * e[2][2])
        z = e[0][2] * (e[1][0] * e[2][1] - e[1][1] * e[2][0])
        det = x + y + z
        inv = np.zeros((3, 3))
        inv[0][0] = e[1][1] * e[2][2] - e[1][2] * e[2][1]
        inv[1][0] = e[1][2] * e[2][0] - e[1][0] * e[2][2]
        inv[2][0] = e[1][0] * e[2][1] - e[1][1] * e[2][0]
        inv[0][1] = e[2][1] * e[0][2] - e[2][2] * e[0][1]
        inv[1][1] = e[2][2] * e[0][0] - e[2][0] * e[0][2]
        inv[2][1] = e[2][0] * e[0][1] - e[2][1] * e[0][0]
        inv[0][2] = e[0][1] * e[1][2] - e[0][2] * e[1][1]
        inv[1][2] = e[0][2] * e[1][0] - e[0][0] * e[1][2]
        inv[2][2] = e[0][0] * e[1][1] - e[0][1] * e[1][0]
        self.frac = inv / det

    def mic(self, d):
        """MinimumImageCartesian for an array of row vectors d[m,3]."""
        f, o = self.frac, self.ortho
        fr = np.stack([(d[:, 0] * f[r][0] + d[:, 1] * f[r][1]) + d[:, 2] * f[r][2] for r in range(3)], axis=1)
        # C round(): half away from zero
        fr = fr - np.copysign(np.floor(np.abs(fr) + 0.5), fr)
        return np.stack([(fr[:, 0] * o[r][0] + fr[:, 1] * o[r][1]) + fr[:, 2] * o[r][2] for r in range(3)], axis=1)


class OBAtom:
    def __init__(self, mol, idx, id_):
        self.mol, self.idx, self.id = mol, idx, id_
        self.z = 0
        self.pos = (0.0, 0.0, 0.0)

    def SetAtomicNum(self, z):
        self.z = int(z)

    def SetVector(self, x, y, z):
        self.pos = (float(x), float(y), float(z))

    def GetId(self):
        return self.id


class OBBond:
    def __init__(self, a, b, order, d2):
        self.a, self.b, self.order, self.d2 = a, b, order, d2

    def GetBeginAtom(self):
        return self.a

    def GetEndAtom(self):
        return self.b

    def GetBondOrder(self):
        return self.order


class OBMol:
    def __init__(self):
        self.atoms = []
        self.bonds = []
        self.uc = None
        self.periodic = False

    def BeginModify(self):
        pass

    def EndModify(self):
        pass

    def NewAtom(self, id_=None):
        a = OBAtom(self, len(self.atoms), len(self.atoms) if id_ is None else id_)
        self.atoms.append(a)
        return a

    def CloneData(self, uc):
        self.uc = uc

    def SetPeriodicMol(self, v=True):
        self.periodic = bool(v)

    def PerceiveBondOrders(self):
        from oracle.perceive import perceive_bond_orders

        n = len(self.atoms)
        if n == 0 or not self.bonds:
            return None
        P = np.array([a.pos for a in self.atoms], dtype=float)
        nbrs = [[] for _ in range(n)]
        for b in self.bonds:  # creation order = every atom's own bond order
            nbrs[b.a.idx].append(b.b.idx)
            nbrs[b.b.idx].append(b.a.idx)
        orders, info = perceive_bond_orders([a.z for a in self.atoms], nbrs,
                                            lambda a, b: tuple(float(x) for x in self._delta(P, a, b)))
        for b in self.bonds:
            b.order = orders[b.a.idx][nbrs[b.a.idx].index(b.b.idx)]
        self.perceive_info = info
        return None

    def _delta(self, P, a, b):
        d = (P[a] - P[b]).reshape(1, 3)
        if self.periodic and self.uc is not None:
            d = self.uc.mic(d)
        return d[0]

    def ConnectTheDots(self):
        n = len(self.atoms)
        if n == 0:
            return
        P = np.array([a.pos for a in self.atoms], dtype=float)
        Zs = np.array([a.z for a in self.atoms], dtype=int)
        order = sorted(range(n), key=lambda i: (P[i, 2], i))
        order = np.array(order, dtype=int)
        rad = np.array([RCOV[z] for z in Zs[order]])
        maxrad = rad.max()
        per = self.periodic and self.uc is not None
        created = []
        for j in range(n - 1):
            ks = np.arange(j + 1, n)
            if not per:
                dz = P[order[j], 2] - P[order[ks], 2]
                maxcut = (rad[j] + maxrad + 0.45) ** 2
                stop = np.nonzero(dz * dz > maxcut)[0]
                if len(stop):
                    ks = ks[: stop[0]]
                if not len(ks):
                    continue
            d = P[order[j]][None, :] - P[order[ks]]
            if per:
                d = self.uc.mic(d)
            d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
            cut = rad[j] + rad[ks] + 0.45
            cut = cut * cut
            ok = ~(d2 > cut) & ~(d2 < 0.16)
            for k, dd in zip(ks[ok], d2[ok]):
                created.append(OBBond(self.atoms[order[j]], self.atoms[order[k]], 1, float(dd)))
        per_atom = [[] for _ in range(n)]
        for b in created:
            per_atom[b.a.idx].append(b)
            per_atom[b.b.idx].append(b)
        dead = set()

        def smallest_angle(i, blist):
            mind = 360.0
            for x in range(len(blist)):
                nb = blist[x].b.idx if blist[x].a.idx == i else blist[x].a.idx
                v1 = self._delta(P, nb, i)
                for y in range(x + 1, len(blist)):
                    nc = blist[y].b.idx if blist[y].a.idx == i else blist[y].a.idx
                    v2 = self._delta(P, nc, i)
                    l1 = math.sqrt((v1[0] * v1[0] + v1[1] * v1[1]) + v1[2] * v1[2])
                    l2 = math.sqrt((v2[0] * v2[0] + v2[1] * v2[1]) + v2[2] * v2[2])
                    if abs(l1) < 1e-3 or abs(l2) < 1e-3:
                        deg = 0.0
                    else:
                        dp = ((v1[0] * v2[0] + v1[1] * v2[1]) + v1[2] * v2[2]) / (l1 * l2)
                        if dp < -0.999999:
                            dp = -0.9999999
                        if dp > 0.9999999:
                            dp = 0.9999999
                        deg = (180.0 / math.pi) * math.acos(dp)
                    if deg < mind:
                        mind = deg
            return mind

        for i in range(n):
            while True:
                blist = [b for b in per_atom[i] if id(b) not in dead]
                if not (len(blist) > MAXBONDS[Zs[i]] or smallest_angle(i, blist) < 45.0):
                    break
                if not blist:
                    break
                if Zs[i] == 1:
                    hit = None
                    for b in blist:
                        o = b.b.idx if b.a.idx == i else b.a.idx
                        if Zs[o] == 1:
                            hit = b
                            break
                    if hit is not None:
                        dead.add(id(hit))
                        continue
                mx = blist[0]
                mxl = math.sqrt(mx.d2)
                for b in blist[1:]:
                    l = math.sqrt(b.d2)
                    if l > mxl:
                        mx, mxl = b, l
                dead.add(id(mx))
        self.bonds = [b for b in created if id(b) not in dead]


def OBMolBondIter(mol):
    return iter(mol.bonds)
