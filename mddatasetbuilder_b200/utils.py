"""Small helpers kept from the reference's utils module.

The reference's ``run_mp`` / ``multiopen`` / ``produce`` Pool fan-out (utils.py:14-96,198-231)
and its lz4+pickle temp-file codec (utils.py:99-195) have no counterpart here: frames are
batched onto the GPU by ``DatasetBuilder`` and intermediate state stays in memory.
"""

from typing import List, TypeVar, Union

T = TypeVar("T")


def must_be_list(obj: Union[T, List[T]]) -> List[T]:
    """Same contract as reference utils.py:244-259: wrap a non-list into a list."""
    if isinstance(obj, list):
        return obj
    return [obj]
