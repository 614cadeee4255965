"""laplace wrapper (dask_image/ndfilters/_diff.py:14-26) -- SURVEY.md 8(f) rank 1."""
import scipy.ndimage

from ..dispatch._dispatch_ndfilters import dispatch_laplace
from . import _utils

__all__ = ["laplace"]


@_utils._update_wrapper(scipy.ndimage.laplace)
def laplace(image, mode='reflect', cval=0.0):
    depth, boundary = _utils._get_depth_boundary(image.ndim, 1, "none")
    return _utils._map_chunks(image, dispatch_laplace, depth, boundary, mode=mode, cval=cval)
