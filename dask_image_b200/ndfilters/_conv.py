"""N-D convolve / correlate wrappers (dask_image/ndfilters/_conv.py:15-66)."""
import scipy.ndimage

from ..dispatch._dispatch_ndfilters import dispatch_convolve, dispatch_correlate
from ..dispatch._utils import check_arraytypes_compatible
from . import _utils

__all__ = ["convolve", "correlate"]


def _nd_filter(dispatcher, image, weights, mode, cval, origin):
    if not hasattr(image, "_run_plan"):
        # a ShardedVolume takes its (small) weights from the host or a B200Array
        check_arraytypes_compatible(image, weights)
    origin = _utils._get_origin(weights.shape, origin)
    depth = _utils._get_depth(weights.shape, origin)
    depth, boundary = _utils._get_depth_boundary(image.ndim, depth, "none")
    if mode == "wrap" and hasattr(image, "map_overlap"):
        # the reference's fix for dask/dask-image#242 (_conv.py:23-25): chunks
        # cannot wrap on their own, so dask supplies periodic halos and the
        # chunk function sees mode="constant".  A whole B200Array / a
        # ShardedVolume wraps natively in the kernel (ring halo exchange).
        boundary = "periodic"
        mode = "constant"
    return _utils._map_chunks(image, dispatcher, depth, boundary,
                              weights=weights, mode=mode, cval=cval, origin=origin)


@_utils._update_wrapper(scipy.ndimage.convolve)
def convolve(image, weights, mode="reflect", cval=0.0, origin=0):
    return _nd_filter(dispatch_convolve, image, weights, mode, cval, origin)


@_utils._update_wrapper(scipy.ndimage.correlate)
def correlate(image, weights, mode="reflect", cval=0.0, origin=0):
    return _nd_filter(dispatch_correlate, image, weights, mode, cval, origin)
