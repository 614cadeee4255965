"""uniform_filter wrapper (dask_image/ndfilters/_smooth.py:17-40)."""
import scipy.ndimage

from ..dispatch._dispatch_ndfilters import dispatch_uniform_filter
from . import _utils

__all__ = ["uniform_filter"]


@_utils._update_wrapper(scipy.ndimage.uniform_filter)
def uniform_filter(image, size=3, mode='reflect', cval=0.0, origin=0):
    size = _utils._get_size(image.ndim, size)
    depth = _utils._get_depth(size, origin)
    depth, boundary = _utils._get_depth_boundary(image.ndim, depth, "none")
    return _utils._map_chunks(image, dispatch_uniform_filter, depth, boundary,
                              size=size, mode=mode, cval=cval, origin=origin)
