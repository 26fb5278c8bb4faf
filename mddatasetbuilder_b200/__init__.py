"""B200-native per-frame analysis path of MDDatasetBuilder.

Drop-in for the ``mddatasetbuilder`` package on this path: ``DatasetBuilder`` keeps the reference's
constructor, ``builddataset`` and on-disk outputs; the work runs in the sm_100a kernels behind
``include/mdb_b200.h`` (``mddatasetbuilder_b200/libmdb_b200.so``).  Nothing here falls back to the CPU.
"""

from ._version import version as __version__
from .datasetbuilder import DatasetBuilder


def get_engine(*args, **kwargs):
    """The process-wide device engine (see ``engine.get_engine``); imported lazily so that importing the
    package does not need CUDA."""
    from .engine import get_engine as _get

    return _get(*args, **kwargs)


__all__ = ["DatasetBuilder", "get_engine", "__version__"]
