"""Seeded synthetic ReaxFF-style C/H/O trajectories.

The reference's test trajectories are downloaded at test time (reference
``tests/test.json:4-15``) and are not available offline, so every benchmark and parity
input is produced here: an orthorhombic periodic box filled with small molecules on a
jittered lattice, ~1 % "reactive contacts" (over-coordinated O, bridging H, squeezed
angles, pairs placed at a bond cutoff +- 1e-6 A) that exercise the valence/angle cleanup
and the borderline compares, and thermal noise per frame.  Coordinates are rounded to
1e-6 A so LAMMPS text round-trips exactly.

Element index convention everywhere in this package: ``types[i]`` is the 0-based index
into ``atomname`` (LAMMPS type - 1), exactly like ``atomname[type-1]`` at reference
``detect.py:87,193,314``.
"""

from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

BASE_SEED = 20261019

# name -> (element symbols, coordinates [A], bonds (i, j, order))
_T = math.acos(-1.0 / 3.0)


def _tetra(r):
    return [(r * math.sqrt(8 / 9), 0.0, -r / 3), (-r * math.sqrt(2 / 9), r * math.sqrt(2 / 3), -r / 3),
            (-r * math.sqrt(2 / 9), -r * math.sqrt(2 / 3), -r / 3), (0.0, 0.0, r)]


def _bent(r, deg):
    h = math.radians(deg) / 2
    return [(r * math.sin(h), r * math.cos(h), 0.0), (-r * math.sin(h), r * math.cos(h), 0.0)]


_CH4 = (["C", "H", "H", "H", "H"], [(0, 0, 0)] + _tetra(1.09), [(0, k, 1.0) for k in range(1, 5)])
_CH3 = (["C", "H", "H", "H"], [(0, 0, 0)] + _tetra(1.09)[:3], [(0, k, 1.0) for k in range(1, 4)])
_H2O = (["O", "H", "H"], [(0, 0, 0)] + _bent(0.96, 104.5), [(0, 1, 1.0), (0, 2, 1.0)])
_O2 = (["O", "O"], [(-0.605, 0, 0), (0.605, 0, 0)], [(0, 1, 2.0)])
_H2 = (["H", "H"], [(-0.37, 0, 0), (0.37, 0, 0)], [(0, 1, 1.0)])
_CO2 = (["C", "O", "O"], [(0, 0, 0), (-1.16, 0, 0), (1.16, 0, 0)], [(0, 1, 2.0), (0, 2, 2.0)])
_CO = (["C", "O"], [(-0.565, 0, 0), (0.565, 0, 0)], [(0, 1, 3.0)])
_OH = (["O", "H"], [(0, 0, 0), (0.97, 0, 0)], [(0, 1, 1.0)])
_H = (["H"], [(0, 0, 0)], [])
_C2H4 = (["C", "C", "H", "H", "H", "H"],
         [(-0.665, 0, 0), (0.665, 0, 0), (-1.23, 0.92, 0), (-1.23, -0.92, 0), (1.23, 0.92, 0), (1.23, -0.92, 0)],
         [(0, 1, 2.0), (0, 2, 1.0), (0, 3, 1.0), (1, 4, 1.0), (1, 5, 1.0)])
_C2H6 = (["C", "C", "H", "H", "H", "H", "H", "H"],
         [(-0.77, 0, 0), (0.77, 0, 0), (-1.16, 1.02, 0), (-1.16, -0.51, 0.88), (-1.16, -0.51, -0.88),
          (1.16, -1.02, 0), (1.16, 0.51, 0.88), (1.16, 0.51, -0.88)],
         [(0, 1, 1.0), (0, 2, 1.0), (0, 3, 1.0), (0, 4, 1.0), (1, 5, 1.0), (1, 6, 1.0), (1, 7, 1.0)])
# --- reactive contacts -------------------------------------------------------
# over-coordinated O (three H; the 1.05 A one is the longest and must be cut)
_H3O = (["O", "H", "H", "H"], [(0, 0, 0), (0.96, 0, 0), (-0.34, 0.91, 0), (-0.35, -0.50, 0.85)],
        [(0, 1, 1.0), (0, 2, 1.0), (0, 3, 1.0)])
# H bridging two O: H has valence 2 > 1 -> the longer O-H goes
_OHO = (["O", "H", "O"], [(-1.15, 0, 0), (0.0, 0, 0), (1.25, 0, 0)], [(0, 1, 0.6), (1, 2, 0.4)])
# squeezed water: H-O-H = 40 deg < 45 deg -> angle rule fires
_H2O_SQ = (["O", "H", "H"], [(0, 0, 0)] + _bent(1.0, 40.0), [(0, 1, 1.0), (0, 2, 0.9)])
# H3 triangle: every H over-coordinated, H-H deletions cascade across atoms
_H3 = (["H", "H", "H"], [(0, 0, 0), (0.9, 0, 0), (0.45, 0.78, 0)], [(0, 1, 0.5), (1, 2, 0.5), (0, 2, 0.5)])
# O...O pairs straddling the 1.77 A cutoff by 1e-6 A
_OO_IN = (["O", "O"], [(-0.8849995, 0, 0), (0.8849995, 0, 0)], [(0, 1, 0.3)])
_OO_OUT = (["O", "O"], [(-0.8850005, 0, 0), (0.8850005, 0, 0)], [])
# H...H pair straddling the 0.4 A lower bound (d^2 < 0.16 is rejected)
_HH_LO = (["H", "H"], [(-0.1999995, 0, 0), (0.1999995, 0, 0)], [])
_HH_LO_IN = (["H", "H"], [(-0.2000005, 0, 0), (0.2000005, 0, 0)], [(0, 1, 1.0)])
# crowded carbon: 5 neighbours, the longest is cut
_CH5 = (["C", "H", "H", "H", "H", "H"], [(0, 0, 0)] + _tetra(1.09) + [(-0.9, -0.3, 0.85)],
        [(0, k, 1.0) for k in range(1, 5)] + [(0, 5, 0.3)])

MOLECULES = {"CH4": _CH4, "CH3": _CH3, "H2O": _H2O, "O2": _O2, "H2": _H2, "CO2": _CO2, "CO": _CO,
             "OH": _OH, "H": _H, "C2H4": _C2H4, "C2H6": _C2H6}
REACTIVE = {"H3O": _H3O, "OHO": _OHO, "H2O_SQ": _H2O_SQ, "H3": _H3, "OO_IN": _OO_IN, "OO_OUT": _OO_OUT,
            "HH_LO": _HH_LO, "HH_LO_IN": _HH_LO_IN, "CH5": _CH5}

_MIX = [("CH4", 0.30), ("O2", 0.25), ("H2O", 0.15), ("CO2", 0.07), ("CO", 0.05), ("H2", 0.05),
        ("OH", 0.03), ("H", 0.02), ("CH3", 0.03), ("C2H4", 0.025), ("C2H6", 0.025)]
_MIX_DENSE = [("C2H6", 0.35), ("C2H4", 0.25), ("CH4", 0.2), ("CH3", 0.05), ("H2", 0.05), ("H2O", 0.05),
              ("CO", 0.03), ("H", 0.02)]


@dataclass
class SynthSystem:
    atomname: list
    types: np.ndarray            # int32 [N], 0-based element index
    pos0: np.ndarray             # float64 [N,3], frame 0
    box: np.ndarray              # float64 [3] orthorhombic lengths (lo = 0)
    rowptr: np.ndarray           # generator topology (bonds-file mode), CSR over atoms
    col: np.ndarray
    bo: np.ndarray               # float64 per CSR entry: order + U(-0.3, 0.3)
    seed: int
    eid: np.ndarray = None       # bond number of every CSR entry (both directions of a bond share it)
    order: np.ndarray = None     # nominal order of every bond
    meta: dict = field(default_factory=dict)

    @property
    def natoms(self):
        return len(self.types)

    def cell(self):
        return np.diag(self.box)


def _random_rotations(rng, m):
    q = rng.normal(size=(m, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    a, b, c, d = q.T
    R = np.empty((m, 3, 3))
    R[:, 0, 0] = a * a + b * b - c * c - d * d
    R[:, 0, 1] = 2 * (b * c - a * d)
    R[:, 0, 2] = 2 * (b * d + a * c)
    R[:, 1, 0] = 2 * (b * c + a * d)
    R[:, 1, 1] = a * a - b * b + c * c - d * d
    R[:, 1, 2] = 2 * (c * d - a * b)
    R[:, 2, 0] = 2 * (b * d - a * c)
    R[:, 2, 1] = 2 * (c * d + a * b)
    R[:, 2, 2] = a * a - b * b - c * c + d * d
    return R


def make_system(natoms: int, density: float, seed: int = BASE_SEED, atomname=("C", "H", "O"),
                reactive_fraction: float = 0.01, dense: bool = False) -> SynthSystem:
    """Build frame 0 of a periodic C/H/O box with exactly ``natoms`` atoms at number
    density ``density`` (atoms / A^3)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    atomname = list(atomname)
    eidx = {s: i for i, s in enumerate(atomname)}
    mix = _MIX_DENSE if dense else _MIX
    lib = dict(MOLECULES)
    lib.update(REACTIVE)
    usable = [(k, w) for k, w in mix if all(s in eidx for s in lib[k][0])]
    react = [k for k in REACTIVE if all(s in eidx for s in REACTIVE[k][0])]
    if not usable:
        raise ValueError("atomname admits no molecule of the library")
    names = [k for k, _ in usable]
    probs = np.array([w for _, w in usable], dtype=float)
    probs /= probs.sum()

    L = (natoms / density) ** (1.0 / 3.0)
    mean_size = sum(len(lib[k][0]) * p for k, p in zip(names, probs))
    nsites_target = max(1, int(math.ceil(natoms / mean_size * 1.25)))
    g = max(1, int(math.ceil(nsites_target ** (1.0 / 3.0))))
    spacing = L / g
    # draw molecules until the atom budget is spent
    want = int(natoms / mean_size * 1.1) + 16
    draw = rng.choice(len(names), size=want, p=probs)
    is_react = rng.random(want) < (reactive_fraction if react else 0.0)
    rdraw = rng.integers(0, max(len(react), 1), size=want)
    chosen = []
    total = 0
    for k in range(want):
        nm = react[rdraw[k]] if is_react[k] else names[draw[k]]
        sz = len(lib[nm][0])
        if total + sz > natoms:
            break
        chosen.append(nm)
        total += sz
    single = "H" if "H" in eidx else atomname[0]
    while total < natoms:
        chosen.append("__single__")
        total += 1
    nmol = len(chosen)
    if nmol > g ** 3:
        g = int(math.ceil(nmol ** (1.0 / 3.0)))
        spacing = L / g
    sites = rng.permutation(g ** 3)[:nmol]
    centers = np.stack([sites % g, (sites // g) % g, sites // (g * g)], axis=1).astype(float)
    jitter = min(0.25, max(0.0, (spacing - 3.4) / 2)) if not dense else 0.1
    centers = (centers + 0.5) * spacing + rng.uniform(-jitter, jitter, size=(nmol, 3))
    R = _random_rotations(rng, nmol)

    types = np.empty(natoms, dtype=np.int32)
    pos = np.empty((natoms, 3))
    edges = []
    w = 0
    counts = {}
    for m, nm in enumerate(chosen):
        if nm == "__single__":
            syms, xyz, bonds = [single], [(0, 0, 0)], []
        else:
            syms, xyz, bonds = lib[nm]
        counts[nm] = counts.get(nm, 0) + 1
        k = len(syms)
        types[w:w + k] = [eidx[s] for s in syms]
        pos[w:w + k] = np.asarray(xyz, dtype=float) @ R[m].T + centers[m]
        for i, j, o in bonds:
            edges.append((w + i, w + j, o))
        w += k
    pos = np.round(np.mod(pos, L), 6)
    pos[pos >= L] -= L
    box = np.array([L, L, L])

    # generator topology as CSR with noisy bond orders (bonds-file mode input)
    deg = np.zeros(natoms + 1, dtype=np.int64)
    for i, j, _ in edges:
        deg[i + 1] += 1
        deg[j + 1] += 1
    rowptr = np.cumsum(deg).astype(np.int32)
    col = np.zeros(int(rowptr[-1]), dtype=np.int32)
    bo = np.zeros(int(rowptr[-1]))
    eid = np.zeros(int(rowptr[-1]), dtype=np.int32)
    fill = rowptr[:-1].astype(np.int64).copy()
    noise = rng.uniform(-0.3, 0.3, size=len(edges))
    for b, ((i, j, o), e) in enumerate(zip(edges, noise)):
        v = round(o + e, 3)
        col[fill[i]] = j
        bo[fill[i]] = v
        eid[fill[i]] = b
        fill[i] += 1
        col[fill[j]] = i
        bo[fill[j]] = v
        eid[fill[j]] = b
        fill[j] += 1
    return SynthSystem(atomname, types, pos, box, rowptr, col, bo, seed,
                       meta={"density": density, "molecules": counts, "spacing": spacing, "dense": dense},
                       eid=eid, order=np.array([o for _, _, o in edges], dtype=np.float64))


def frame_positions(sys: SynthSystem, nframes: int, sigma: float = 0.05, seed: int | None = None) -> np.ndarray:
    """Host frames [F, N, 3]: frame 0 is exact, frame f>0 adds N(0, sigma) and wraps."""
    rng = np.random.Generator(np.random.PCG64((sys.seed if seed is None else seed) + 7919))
    out = np.empty((nframes, sys.natoms, 3))
    out[0] = sys.pos0
    for f in range(1, nframes):
        p = sys.pos0 + rng.normal(0.0, sigma, size=sys.pos0.shape)
        p = np.round(np.mod(p, sys.box), 6)
        p[p >= sys.box] -= sys.box[0]
        out[f] = p
    return out


def frame_positions_device(sys: SynthSystem, nframes: int, device, sigma: float = 0.05, seed: int | None = None):
    """Device frames [F, N, 3] float64 drawn with torch's counter-based Philox generator
    (frame 0 exact).  Used for the full-size benchmark shapes, which are too large to
    draw on the host."""
    import torch

    gen = torch.Generator(device=device)
    gen.manual_seed((sys.seed if seed is None else seed) + 104729)
    base = torch.as_tensor(sys.pos0, device=device)
    box = torch.as_tensor(sys.box, device=device)
    out = torch.empty((nframes, sys.natoms, 3), dtype=torch.float64, device=device)
    chunk = max(1, int(2 ** 26 // max(sys.natoms, 1)))
    for f0 in range(0, nframes, chunk):
        f1 = min(nframes, f0 + chunk)
        noise = torch.randn((f1 - f0, sys.natoms, 3), dtype=torch.float32, device=device, generator=gen)
        p = base[None] + sigma * noise.double()
        p = torch.round(torch.remainder(p, box) * 1e6) / 1e6
        p = torch.where(p >= box, p - box, p)
        out[f0:f1] = p
    out[0] = base
    return out


def bond_table_device(sys: SynthSystem, nframes: int, device, seed: int | None = None, amplitude: float = 0.3):
    """Bond-file mode input for `nframes` frames on the device, as CSR: rowptr [F, N+1] int32 (frame-local),
    col [F, nnz] int32 and bo [F, nnz] float64.  The topology is the generator's; every frame draws its own
    bond orders `nominal + U(-amplitude, amplitude)` per bond (both directions of a bond carry the same value,
    like a `fix reaxff/bonds` table), rounded to the 3 decimals the text format keeps.  Frame 0 = `sys.bo`."""
    import torch

    gen = torch.Generator(device=device)
    gen.manual_seed((sys.seed if seed is None else seed) + 15485863)
    nnz = int(sys.rowptr[-1])
    rowptr = torch.as_tensor(sys.rowptr, dtype=torch.int32, device=device)[None].repeat(nframes, 1)
    col = torch.as_tensor(sys.col, dtype=torch.int32, device=device)[None].repeat(nframes, 1)
    eid = torch.as_tensor(sys.eid, dtype=torch.int64, device=device)
    nominal = torch.as_tensor(sys.order, dtype=torch.float64, device=device)
    bo = torch.empty((nframes, nnz), dtype=torch.float64, device=device)
    chunk = max(1, int(2 ** 26 // max(len(sys.order), 1)))
    for f0 in range(0, nframes, chunk):
        f1 = min(nframes, f0 + chunk)
        u = torch.rand((f1 - f0, len(sys.order)), dtype=torch.float32, device=device, generator=gen).double()
        v = torch.round((nominal[None] + amplitude * (2.0 * u - 1.0)) * 1e3) / 1e3
        bo[f0:f1] = v[:, eid]
    bo[0] = torch.as_tensor(sys.bo, dtype=torch.float64, device=device)
    return rowptr.contiguous(), col.contiguous(), bo


# ----------------------------------------------------------------------------
# LAMMPS text (formats: SURVEY.md App. B; parsed at reference detect.py:56-88,151-194,294-345)
# ----------------------------------------------------------------------------
def write_lammps_dump(path, sys: SynthSystem, frames: np.ndarray, timestep0=0, dt=1, shuffle_seed=None,
                      cols=("id", "type", "x", "y", "z")):
    rng = np.random.Generator(np.random.PCG64(shuffle_seed)) if shuffle_seed is not None else None
    n = sys.natoms
    with open(path, "w") as f:
        for k, P in enumerate(frames):
            f.write("ITEM: TIMESTEP\n%d\n" % (timestep0 + k * dt))
            f.write("ITEM: NUMBER OF ATOMS\n%d\n" % n)
            f.write("ITEM: BOX BOUNDS pp pp pp\n")
            for c in range(3):
                f.write("%.6f %.6f\n" % (0.0, sys.box[c]))
            f.write("ITEM: ATOMS " + " ".join(cols) + "\n")
            order = rng.permutation(n) if rng is not None else range(n)
            for i in order:
                vals = {"id": str(i + 1), "type": str(int(sys.types[i]) + 1), "x": "%.6f" % P[i, 0],
                        "y": "%.6f" % P[i, 1], "z": "%.6f" % P[i, 2], "q": "0.0"}
                f.write(" ".join(vals[c] for c in cols) + "\n")


def write_lammps_bonds(path, sys: SynthSystem, nframes: int, timestep0=0, dt=1, bo_frames=None):
    """``fix reaxff/bonds`` text: 7 header lines, N atom lines, one closing '#'."""
    n = sys.natoms
    with open(path, "w") as f:
        for k in range(nframes):
            bo = sys.bo if bo_frames is None else bo_frames[k]
            f.write("# Timestep %d\n#\n# Number of particles %d\n#\n" % (timestep0 + k * dt, n))
            f.write("# Max number of bonds per atom 4 with coarse bond order cutoff 0.300\n")
            f.write("# Particle connection table and bond orders\n")
            f.write("# id type nb id_1...id_nb mol bo_1...bo_nb abo nlp q\n")
            for i in range(n):
                a, b = sys.rowptr[i], sys.rowptr[i + 1]
                nb = b - a
                ids = " ".join(str(int(j) + 1) for j in sys.col[a:b])
                bos = " ".join("%.3f" % v for v in bo[a:b])
                f.write(" %d %d %d %s 0 %s %.3f 0.000 0.000\n" % (i + 1, int(sys.types[i]) + 1, nb, ids, bos,
                                                                  float(np.sum(bo[a:b]))))
            f.write("#\n")
