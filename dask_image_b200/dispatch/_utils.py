"""Array-type compatibility check (dask_image/dispatch/_utils.py:10-25)."""
from ._dispatcher import get_type

__all__ = ["check_arraytypes_compatible"]


def check_arraytypes_compatible(*args):
    """Raise ``ValueError`` unless every argument is backed by the same array
    type (a dask array counts as the type of its chunks)."""
    kinds = {get_type(a) for a in args}
    if len(kinds) != 1:
        raise ValueError("Array types must be compatible.")
