"""A stand-in for the sliver of dask that dask-image's ndfilters path and its own tests touch -- TEST INFRASTRUCTURE.

dask cannot be installed in this image (no network, not in the wheelhouse).  The reference's ndfilters code imports
one name from it, ``dask.utils.Dispatch`` (ref:dask_image/dispatch/_dispatcher.py:3), and calls ``map_overlap`` on
the array it is given; the reference's ndfilters TESTS additionally use ``dask.array.from_array``, ``stack``,
``concatenate`` and ``dask.array.utils.assert_eq``.  This package provides exactly those, eagerly, on top of
``dask_image_b200.HaloChunked`` (chunks are ``B200Array`` objects), so that the UNMODIFIED reference wrappers -- and
the reference's own test files -- run against this repository's chunk type (tests/test_reference_suite_cpu.py).
It is never importable from the product: only test runners put tests/_dask_stub on sys.path.
"""
from . import utils  # noqa: F401

__version__ = "0+b200.stub"
