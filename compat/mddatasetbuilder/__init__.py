"""Alias of ``mddatasetbuilder_b200`` under the reference's import name, so that
``from mddatasetbuilder import DatasetBuilder`` and ``python -m mddatasetbuilder`` (reference
``__init__.py:4-6``, ``__main__.py:3-6``, ``tests/test_cli.py:9``) keep working after the switch.
Installed only through pyproject.toml; it is not on the path of an in-tree checkout, where the reference's
own package may be importable for the golden runs (oracle/refrun.py)."""

from mddatasetbuilder_b200 import DatasetBuilder, __version__  # noqa: F401
from mddatasetbuilder_b200 import datasetbuilder, detect, dps, utils  # noqa: F401

__all__ = ["DatasetBuilder"]
