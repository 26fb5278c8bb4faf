# B200-native per-frame analysis path of MDDatasetBuilder (drop-in for `mddatasetbuilder`).
"""MDDatasetBuilder."""

from .datasetbuilder import DatasetBuilder

__all__ = ["DatasetBuilder"]
