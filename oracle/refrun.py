"""Import the reference's own, unmodified Python from /root/reference through the shims.

TEST INFRASTRUCTURE ONLY, and only usable in the build container (where /root/reference is
mounted): used by ``tests/golden/make_golden.py`` to generate committed fixtures and by the
not-gpu tests that pin the oracle against the reference when the tree is present.  Nothing is
copied: the package is assembled in ``sys.modules`` with ``__path__`` spanning the reference
directory and ``oracle/_ref/mddatasetbuilder`` (generated ``_version.py`` + compiled ``dps``).
"""
import importlib
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("MDB_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "mddatasetbuilder")) and os.path.isdir(
        os.path.join(_HERE, "_ref", "mddatasetbuilder"))


def load():
    """Return the reference package (modules detect, datasetbuilder, utils usable)."""
    if "mddatasetbuilder" in sys.modules and getattr(sys.modules["mddatasetbuilder"], "_mdb_ref", False):
        return sys.modules["mddatasetbuilder"]
    if not available():
        raise RuntimeError("reference tree or oracle/_ref missing (run `make -C oracle`)")
    shims = os.path.join(_HERE, "shims")
    if shims not in sys.path:
        sys.path.insert(0, shims)
    pkg = types.ModuleType("mddatasetbuilder")
    pkg.__path__ = [os.path.join(REF, "mddatasetbuilder"), os.path.join(_HERE, "_ref", "mddatasetbuilder")]
    pkg._mdb_ref = True
    sys.modules["mddatasetbuilder"] = pkg
    for name in ("_logger", "utils", "detect", "datasetbuilder"):
        setattr(pkg, name, importlib.import_module(f"mddatasetbuilder.{name}"))
    pkg.DatasetBuilder = pkg.datasetbuilder.DatasetBuilder
    return pkg
