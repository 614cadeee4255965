"""minimum_filter / maximum_filter wrappers (dask_image/ndfilters/_order.py:21-99)
-- SURVEY.md 8(f) rank 4.  The median / rank / percentile filters of that module
are rank statistics, not correlations, and stay outside the B200 path."""
import scipy.ndimage

from ..dispatch._dispatch_ndfilters import dispatch_maximum_filter, dispatch_minimum_filter
from . import _utils

__all__ = ["minimum_filter", "maximum_filter"]


def _extreme(dispatcher, image, size, footprint, mode, cval, origin):
    footprint = _utils._get_footprint(image.ndim, size, footprint)
    origin = _utils._get_origin(footprint.shape, origin)
    depth = _utils._get_depth(footprint.shape, origin)
    depth, boundary = _utils._get_depth_boundary(footprint.ndim, depth, "none")
    return _utils._map_chunks(image, dispatcher, depth, boundary,
                              footprint=footprint, mode=mode, cval=cval, origin=origin)


@_utils._update_wrapper(scipy.ndimage.minimum_filter)
def minimum_filter(image, size=None, footprint=None, mode='reflect', cval=0.0, origin=0):
    return _extreme(dispatch_minimum_filter, image, size, footprint, mode, cval, origin)


@_utils._update_wrapper(scipy.ndimage.maximum_filter)
def maximum_filter(image, size=None, footprint=None, mode='reflect', cval=0.0, origin=0):
    return _extreme(dispatch_maximum_filter, image, size, footprint, mode, cval, origin)
