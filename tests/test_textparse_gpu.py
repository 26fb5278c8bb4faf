"""GPU parity: the device-side LAMMPS dump parser (csrc/textparse.cu) against the host parser
(csrc/textio.cpp, std::from_chars = the value Python's float() gives, reference detect.py:294-345).
Bit-exact positions and cells, whatever the number format; frames outside the exact fast path are
handed back to the host parser."""
import os

import numpy as np
import pytest

from mddatasetbuilder_b200.reader import KIND_DUMP, TextReader

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
DUMP = os.path.join(HERE, "golden", "traj", "dump.small")


def _read_all_host(path, stride=1, batch=3):
    r = TextReader(KIND_DUMP, path, nthreads=2)
    P, Cc = [], []
    while True:
        pos, cell = r.read(batch, stride)
        if len(pos) == 0:
            break
        P.append(pos.copy())
        Cc.append(cell.copy())
    r.close()
    return np.concatenate(P), np.concatenate(Cc)


def _read_all_device(engine, path, stride=1, batch=3):
    r = TextReader(KIND_DUMP, path, nthreads=2)
    P, Cc = [], []
    while True:
        pos, cell = r.read_device(batch, engine, stride)
        if pos.shape[0] == 0:
            break
        P.append(pos.cpu().numpy())
        Cc.append(cell.copy())
    fb = getattr(r, "device_fallback_frames", 0)
    r.close()
    return np.concatenate(P), np.concatenate(Cc), fb


def _same_bits(a, b):
    return a.shape == b.shape and np.array_equal(a.view(np.uint64), b.view(np.uint64))


@pytest.mark.parametrize("stride,batch", [(1, 3), (2, 4), (5, 1)])
def test_golden_dump_bit_exact(engine, stride, batch):
    hp, hc = _read_all_host(DUMP, stride, batch)
    dp, dc, fb = _read_all_device(engine, DUMP, stride, batch)
    assert _same_bits(hp, dp) and _same_bits(hc, dc)
    assert fb == 0, "the reference trajectory must stay on the device fast path"


def _write_dump(path, ids, types, frames, fmt, cols=("id", "type", "x", "y", "z"), sep=" ", eol="\n", tilt=None,
                plus=False):
    with open(path, "w", newline="") as f:
        for k, P in enumerate(frames):
            f.write("ITEM: TIMESTEP" + eol + str(100 * k) + eol)
            f.write("ITEM: NUMBER OF ATOMS" + eol + str(len(ids)) + eol)
            if tilt is None:
                f.write("ITEM: BOX BOUNDS pp pp pp" + eol)
                for c in range(3):
                    f.write("%.10g %.10g" % (-1.5 * c, 20.0 + c) + eol)
            else:
                f.write("ITEM: BOX BOUNDS xy xz yz pp pp pp" + eol)
                for c in range(3):
                    f.write("%.10g %.10g %.10g" % (-2.0, 22.0 + c, tilt[c]) + eol)
            f.write("ITEM: ATOMS " + " ".join(cols) + eol)
            for i in np.random.default_rng(k).permutation(len(ids)):
                def num(v):
                    s = fmt % v
                    return "+" + s if plus and v > 0 else s
                vals = {"id": str(ids[i]), "type": str(types[i]), "x": num(P[i, 0]), "y": num(P[i, 1]), "z": num(P[i, 2]),
                        "q": "-0.25", "vx": "1e-3"}
                f.write(sep.join(vals[c] for c in cols) + eol)


FORMATS = [
    ("%.6f", {}, 0),
    ("%g", {}, 0),
    ("%.15g", {}, 0),
    ("%.8e", {}, 0),
    ("%.3f", dict(cols=("type", "q", "z", "id", "vx", "y", "x"), sep="\t"), 0),
    ("%.10f", dict(eol="\r\n", plus=True), 0),
    ("%.12g", dict(tilt=(1.25, -0.5, 0.75)), 0),
    ("%.17g", {}, None),   # 17 significant digits exceed 2^53: those frames go to the host parser
    ("%.25f", {}, None),   # more than 19 digits
]


@pytest.mark.parametrize("fmt,kw,fallback", FORMATS)
def test_number_formats_bit_exact(engine, tmp_path, fmt, kw, fallback):
    rng = np.random.default_rng(abs(hash(fmt)) % (1 << 31))
    n = 1237
    ids = np.arange(1, n + 1)
    types = rng.integers(1, 4, n)
    frames = rng.normal(0.0, 30.0, (5, n, 3))
    frames[0, :7] = [[0.0, -0.0, 1e-12], [1e22, -1e-22, 123456789012345.0], [9007199254740992.0, 0.5, -0.125],
                     [1e-5, 2.5e-7, 3.0], [1.0, 10.0, 100.0], [0.1, 0.2, 0.3], [-7.25e10, 6.02e17, 1.6e-19]]
    path = str(tmp_path / "dump.fmt")
    _write_dump(path, ids, types, frames, fmt, **kw)
    hp, hc = _read_all_host(path, 1, 2)
    dp, dc, fb = _read_all_device(engine, path, 1, 2)
    assert _same_bits(hp, dp), "positions differ from the host parser"
    assert _same_bits(hc, dc)
    if fallback is not None:
        assert fb == fallback
    else:
        assert fb > 0


def test_malformed_frame_raises_like_the_host_parser(engine, tmp_path):
    n = 50
    ids = np.arange(1, n + 1)
    types = np.ones(n, dtype=int)
    frames = np.random.default_rng(1).random((3, n, 3))
    path = str(tmp_path / "dump.bad")
    _write_dump(path, ids, types, frames, "%.6f")
    text = open(path).read().split("\n")
    # second frame, some atom line: a token that is not a number
    k = (9 + n) + 9 + 7
    parts = text[k].split()
    parts[3] = "oops"
    text[k] = " ".join(parts)
    open(path, "w").write("\n".join(text))
    r = TextReader(KIND_DUMP, path)
    with pytest.raises(RuntimeError, match="malformed"):
        r.read_device(3, engine)
    r.close()


def test_batches_feed_the_bond_path(engine):
    """Text -> device positions -> bonds, against host-parsed positions through the same engine."""
    r1 = TextReader(KIND_DUMP, DUMP)
    r2 = TextReader(KIND_DUMP, DUMP)
    pos_h, cell_h = r1.read(4)
    pos_d, cell_d = r2.read_device(4, engine)
    types = (r1.atomtype - 1).astype(np.uint8)
    engine.set_elements(["H", "O"])
    try:
        a = engine.analyze(pos_h, cell_h, types)
        b = engine.analyze(pos_d, cell_d, types)
    finally:
        engine.set_elements(("C", "H", "O"))  # the fixture is shared
    for key in ("ell", "keys", "labels"):
        assert np.array_equal(a[key].cpu().numpy(), b[key].cpu().numpy())
    r1.close()
    r2.close()


def test_prefetch_double_buffering(engine):
    """prefetch=True stages the next batch on a background thread: same frames, and the iteration
    methods refuse to run while a staged batch is pending."""
    hp, hc = _read_all_host(DUMP, 1, 3)
    r = TextReader(KIND_DUMP, DUMP)
    P = []
    pos, cell = r.read_device(3, engine, prefetch=True)
    with pytest.raises(RuntimeError, match="pending"):
        r.skip(1)
    while pos.shape[0]:
        P.append(pos.cpu().numpy())
        pos, cell = r.read_device(3, engine, prefetch=True)
    assert _same_bits(np.concatenate(P), hp)
    r.rewind()
    pos, _ = r.read_device(2, engine, prefetch=True)
    r.rewind()  # drops the staged batch
    pos2, _ = r.read_device(2, engine)
    assert _same_bits(pos.cpu().numpy(), pos2.cpu().numpy())
    r.close()


# --------------------------------------------------------------------------------------------------
# fix reaxff/bonds tables
# --------------------------------------------------------------------------------------------------
BONDS = os.path.join(HERE, "golden", "traj", "bonds.small")


def _bonds_host(path, stride, batch, width=8):
    from mddatasetbuilder_b200.reader import KIND_BOND

    r = TextReader(KIND_BOND, path, nthreads=2)
    A, B = [], []
    while True:
        nbr, bo = r.read(batch, stride, width=width)
        if len(nbr) == 0:
            break
        A.append(nbr.copy())
        B.append(bo.copy())
    r.close()
    return np.concatenate(A), np.concatenate(B)


def _bonds_device(engine, path, stride, batch, width=8, prefetch=False):
    from mddatasetbuilder_b200.reader import KIND_BOND

    r = TextReader(KIND_BOND, path, nthreads=2)
    A, B = [], []
    while True:
        nbr, bo = r.read_device(batch, engine, stride, prefetch=prefetch, width=width)
        if nbr.shape[0] == 0:
            break
        A.append(nbr.cpu().numpy())
        B.append(bo.cpu().numpy())
    fb = getattr(r, "device_fallback_frames", 0)
    r.close()
    return np.concatenate(A), np.concatenate(B), fb


@pytest.mark.parametrize("stride,batch,prefetch", [(1, 3, False), (2, 5, True), (7, 1, False)])
def test_golden_bonds_table_bit_exact(engine, stride, batch, prefetch):
    hn, hb = _bonds_host(BONDS, stride, batch)
    dn, db, fb = _bonds_device(engine, BONDS, stride, batch, prefetch=prefetch)
    assert np.array_equal(hn, dn) and _same_bits(hb, db)
    assert fb == 0


def test_synthetic_bonds_table_and_errors(engine, tmp_path):
    from mddatasetbuilder_b200 import synth
    from mddatasetbuilder_b200.reader import KIND_BOND

    s = synth.make_system(4000, 0.05, seed=synth.BASE_SEED + 61)
    path = str(tmp_path / "bonds.syn")
    rng = np.random.default_rng(5)
    bo_frames = [np.abs(s.bo + rng.uniform(-0.3, 0.3, s.bo.shape)) for _ in range(4)]
    synth.write_lammps_bonds(path, s, 4, bo_frames=bo_frames)
    hn, hb = _bonds_host(path, 1, 3, width=6)
    dn, db, fb = _bonds_device(engine, path, 1, 3, width=6)
    assert np.array_equal(hn, dn) and _same_bits(hb, db) and fb == 0
    # a row wider than `width`: the device parser flags the frame, the host parser names the problem
    r = TextReader(KIND_BOND, path)
    with pytest.raises(RuntimeError, match="raise max_bonds"):
        r.read_device(2, engine, width=2)
    r.close()
