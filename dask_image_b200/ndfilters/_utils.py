"""Argument normalisation and halo sizing shared by the ndfilters wrappers.

Behavioural mirror of dask_image/ndfilters/_utils.py (exception classes are
pinned by the reference's tests/test_dask_image/test_ndfilters/test__utils.py:63-128,
test__conv.py:19-43 and test__smooth.py:12-28); written from that contract.
"""
import inspect
import numbers
import re
from collections import abc

import numpy as np

_STR = (str, bytes)


def _get_docstring(func):
    """Docstring of ``func`` with the ``output`` parameter dropped, ``input``
    renamed to ``image`` and the examples cut (reference: _utils.py:10-39)."""
    text = func.__doc__ or ""
    # split in front of every 4-space-indented "name : " parameter header
    pieces = re.sub(r"(    [A-Za-z]+ : )", "\0\\1", text).split("\0")
    text = "".join(p for p in pieces if not p.startswith("    output : "))
    text = text.replace("input", "image").replace("labels", "label_image")
    text = text.split("Examples")[0].strip()
    template = (
        "\n    Wrapped copy of \"{mod_name}.{func_name}\"\n\n\n"
        "    Excludes the output parameter as it would not work with Dask arrays.\n\n\n"
        "    Original docstring:\n\n    {doc}\n    "
    )
    return template.format(mod_name=inspect.getmodule(func).__name__,
                           func_name=func.__name__, doc=text)


def _update_wrapper(func):
    """Decorator: take name and (cleaned) docstring from ``func``."""
    def deco(wrapper):
        wrapper.__name__ = func.__name__
        wrapper.__doc__ = _get_docstring(func)
        return wrapper

    return deco


def _per_axis_mapping(value, ndim, what, scalar_ok):
    """scalar -> tuple; sized -> must have ndim entries; sequence -> {axis: v}."""
    if scalar_ok(value):
        value = (value,) * ndim
    if not isinstance(value, abc.Sized):
        raise TypeError("Unexpected type for `%s`." % what)
    if len(value) != ndim:
        raise ValueError("Expected `%s` to have a length equal to `ndim`." % what)
    if isinstance(value, abc.Sequence):
        value = {ax: v for ax, v in enumerate(value)}
    if not isinstance(value, abc.Mapping):
        raise TypeError("Unexpected type for `%s`." % what)
    return value


def _get_depth_boundary(ndim, depth, boundary=None):
    """Per-axis ``{axis: depth}`` / ``{axis: boundary}`` dicts for ``map_overlap``."""
    if not isinstance(ndim, numbers.Integral):
        raise TypeError("Expected integer value for `ndim`.")
    if ndim <= 0:
        raise ValueError("Expected positive value for `ndim`.")

    depth = _per_axis_mapping(depth, ndim, "depth", lambda v: isinstance(v, numbers.Number))
    for d in depth.values():
        if not isinstance(d, numbers.Integral):
            raise TypeError("Expected integer values for `depth`.")
    for d in depth.values():
        if d < 0:
            raise ValueError("Expected positive semidefinite values for `depth`.")
    depth = {ax: int(d) for ax, d in depth.items()}

    boundary = _per_axis_mapping(boundary, ndim, "boundary",
                                 lambda v: v is None or isinstance(v, _STR))
    for b in boundary.values():
        if not (b is None or isinstance(b, _STR)):
            raise TypeError("Expected string-like values for `boundary`.")
    return depth, boundary


def _get_size(ndim, size):
    """Filter extent per axis as a tuple of ints."""
    if not isinstance(ndim, numbers.Integral):
        raise TypeError("The ndim must be of integral type.")
    if isinstance(size, numbers.Number):
        size = (size,) * ndim
    size = np.array(size)
    if size.ndim != 1:
        raise RuntimeError("The size must have only one dimension.")
    if len(size) != ndim:
        raise RuntimeError("The size must have a length equal to the number of dimensions.")
    if not issubclass(size.dtype.type, numbers.Integral):
        raise TypeError("The size must be of integral type.")
    return tuple(size)


def _get_origin(size, origin=0):
    """Origin per axis, validated against the footprint ``size``."""
    size = np.array(size)
    ndim = len(size)
    if isinstance(origin, numbers.Number):
        origin = (origin,) * ndim
    origin = np.array(origin)
    if not issubclass(origin.dtype.type, numbers.Integral):
        raise TypeError("The origin must be of integral type.")
    if origin.ndim != 1:
        raise RuntimeError("The origin must have only one dimension.")
    if len(origin) != ndim:
        raise RuntimeError("The origin must have the same length as the number of dimensions"
                           " as the array being filtered.")
    if not np.all(origin < (size + 1) // 2):
        raise ValueError("The origin must be within the footprint.")
    return tuple(origin)


def _get_depth(size, origin=0):
    """Halo needed per axis: half the footprint plus the origin shift."""
    origin = np.array(_get_origin(size, origin))
    return tuple(np.array(size) // 2 + abs(origin))


def _get_footprint(ndim, size=None, footprint=None):
    """Boolean footprint from ``size`` or a given array (not used by the
    correlation path; kept because the wrappers' helper surface is shared)."""
    if size is None and footprint is None:
        raise RuntimeError("Must provide either size or footprint.")
    if size is not None and footprint is not None:
        raise RuntimeError("Provide either size or footprint, but not both.")
    if size is not None:
        footprint = np.ones(_get_size(ndim, size), dtype=bool)
    if footprint.ndim != ndim:
        raise RuntimeError("The footprint must have the same number of dimensions as"
                           " the array being filtered.")
    if footprint.size == 0:
        raise RuntimeError("The footprint must have only non-zero dimensions.")
    return footprint != 0


def _map_chunks(image, dispatcher, depth, boundary, **kwargs):
    """Run the per-chunk function chosen by ``dispatcher`` over ``image``.

    * dask arrays (anything with ``map_overlap``): exactly the reference's call,
      e.g. dask_image/ndfilters/_gaussian.py:70-81 -- dask assembles halos.
    * ``ShardedVolume``: the dispatcher returns the slab-engine function
      (sharded.py), which exchanges axis-0 halo planes over NCCL instead.
    * a bare ``B200Array`` is one chunk: the chunk function is applied directly
      (``map_overlap`` on a single chunk with ``boundary="none"`` adds nothing).
    """
    if hasattr(image, "map_overlap"):
        return image.map_overlap(dispatcher(image), depth=depth, boundary=boundary,
                                 dtype=image.dtype, meta=image._meta, **kwargs)
    return dispatcher(image)(image, **kwargs)
