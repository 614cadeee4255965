"""Gaussian family wrappers with the reference's signatures
(dask_image/ndfilters/_gaussian.py:58-154)."""
import numbers

import numpy as np
import scipy.ndimage

from ..dispatch._dispatch_ndfilters import (dispatch_gaussian_filter,
                                            dispatch_gaussian_gradient_magnitude,
                                            dispatch_gaussian_laplace)
from . import _utils
from ..ndimage import _freeze

__all__ = ["gaussian_filter", "gaussian_gradient_magnitude", "gaussian_laplace", "gaussian"]


def _get_sigmas(image, sigma):
    """One real sigma per image axis (reference: _gaussian.py:22-44)."""
    arr = np.array(sigma)
    if arr.ndim == 0:
        arr = np.array([arr[()]] * image.ndim)
    if arr.ndim != 1:
        raise RuntimeError("Must have a single sigma or a single sequence.")
    if len(arr) != image.ndim:
        raise RuntimeError("Must have an equal number of sigmas to image dimensions.")
    if not issubclass(arr.dtype.type, numbers.Real):
        raise TypeError("Must have real sigmas.")
    return tuple(arr)


def _get_border(image, sigma, truncate):
    """Halo = ceil(sigma * truncate) per axis (reference: _gaussian.py:47-55)."""
    sigmas = np.array(_get_sigmas(image, sigma))
    if not isinstance(truncate, numbers.Real):
        raise TypeError("Must have a real truncate value.")
    return tuple(np.ceil(sigmas * truncate).astype(int))


_halo_cache = {}


def _halo(image, sigma, truncate):
    """(sigmas, depth, boundary) of a call.  Depends on (ndim, sigma, truncate) only: chunk-wise
    callers repeat the same call per chunk, so validated results are remembered (errors are not)."""
    try:
        key = (image.ndim, _freeze(sigma), _freeze(truncate))
        hit = _halo_cache.get(key)
        if hit is not None:
            return hit[0], dict(hit[1]), dict(hit[2])
    except TypeError:
        key = None
    sigma = _get_sigmas(image, sigma)
    depth = _get_border(image, sigma, truncate)
    depth, boundary = _utils._get_depth_boundary(image.ndim, depth, "none")
    if key is not None:
        if len(_halo_cache) >= 256:
            _halo_cache.clear()
        _halo_cache[key] = (sigma, depth, boundary)
    return sigma, depth, boundary


@_utils._update_wrapper(scipy.ndimage.gaussian_filter)
def gaussian_filter(image, sigma, order=0, mode='reflect', cval=0.0, truncate=4.0):
    sigma, depth, boundary = _halo(image, sigma, truncate)
    return _utils._map_chunks(image, dispatch_gaussian_filter, depth, boundary,
                              sigma=sigma, order=order, mode=mode, cval=cval, truncate=truncate)


def gaussian(image, sigma, order=0, mode='reflect', cval=0.0, truncate=4.0):
    """Alias of `dask_image.ndfilters.gaussian_filter`."""
    return gaussian_filter(image, sigma, order=order, mode=mode, cval=cval, truncate=truncate)


@_utils._update_wrapper(scipy.ndimage.gaussian_gradient_magnitude)
def gaussian_gradient_magnitude(image, sigma, mode='reflect', cval=0.0, truncate=4.0, **kwargs):
    sigma, depth, boundary = _halo(image, sigma, truncate)
    return _utils._map_chunks(image, dispatch_gaussian_gradient_magnitude, depth, boundary,
                              sigma=sigma, mode=mode, cval=cval, truncate=truncate, **kwargs)


@_utils._update_wrapper(scipy.ndimage.gaussian_laplace)
def gaussian_laplace(image, sigma, mode='reflect', cval=0.0, truncate=4.0, **kwargs):
    sigma, depth, boundary = _halo(image, sigma, truncate)
    return _utils._map_chunks(image, dispatch_gaussian_laplace, depth, boundary,
                              sigma=sigma, mode=mode, cval=cval, truncate=truncate, **kwargs)
