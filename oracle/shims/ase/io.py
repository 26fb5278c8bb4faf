"""ase.io stand-in: write(..., format='xyz') = ASE's simple xyz writer. Test infrastructure only."""


def write(filename, images, format=None, **kw):
    assert format in (None, "xyz")
    atoms = images
    with open(filename, "w") as f:
        f.write("%d\n%s\n" % (len(atoms), ""))
        for s, (x, y, z) in zip(atoms.get_chemical_symbols(), atoms.positions):
            f.write("%-2s %s %s %s\n" % (s, "%22.15f" % x, "%22.15f" % y, "%22.15f" % z))
