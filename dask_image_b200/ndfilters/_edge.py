"""prewitt / sobel wrappers (dask_image/ndfilters/_edge.py:17-57) -- SURVEY.md 8(f) rank 1."""
import numbers

import scipy.ndimage

from ..dispatch._dispatch_ndfilters import dispatch_prewitt, dispatch_sobel
from . import _utils

__all__ = ["prewitt", "sobel"]


def _validate_axis(ndim, axis):
    if not isinstance(axis, numbers.Integral):
        raise ValueError("The axis must be of integral type.")
    if not -ndim <= axis < ndim:
        raise ValueError("The axis is out of range.")


def _edge(dispatcher, image, axis, mode, cval):
    _validate_axis(image.ndim, axis)
    depth, boundary = _utils._get_depth_boundary(image.ndim, 1, "none")
    return _utils._map_chunks(image, dispatcher, depth, boundary, axis=axis, mode=mode, cval=cval)


@_utils._update_wrapper(scipy.ndimage.prewitt)
def prewitt(image, axis=-1, mode='reflect', cval=0.0):
    return _edge(dispatch_prewitt, image, axis, mode, cval)


@_utils._update_wrapper(scipy.ndimage.sobel)
def sobel(image, axis=-1, mode='reflect', cval=0.0):
    return _edge(dispatch_sobel, image, axis, mode, cval)
