"""threshold_local (dask_image/ndfilters/_threshold.py:12-97) -- SURVEY.md 8(f)
rank 2: a direct CALLER of the hot path.  ``method='gaussian'`` is a float64
``gaussian_filter``; ``method='mean'`` is the box mean the reference obtains
from ``generic_filter(np.mean)`` and is computed here by ``uniform_filter`` (the
same window, equal to rounding).  ``'generic'`` and ``'median'`` are not part of
the correlation path (arbitrary Python callables / rank filters) and raise."""
import numpy as np

from . import _gaussian, _smooth
from ._unsupported import OutOfScopeError

__all__ = ["threshold_local"]


def threshold_local(image, block_size, method='gaussian', offset=0, mode='reflect', param=None, cval=0):
    """Compute a threshold mask image based on local pixel neighborhood
    (signature and semantics of ``dask_image.ndfilters.threshold_local``).

    Returns the threshold image (float64); pixels of ``image`` above it are
    foreground."""
    image = image.astype(np.float64)
    if method == 'generic':
        if not callable(param):
            raise ValueError("Must include a valid function to use as the 'param' keyword argument.")
        raise OutOfScopeError("method='generic' runs an arbitrary Python callable per voxel; "
                              "it is outside the B200 correlation path")
    elif method == 'gaussian':
        if param is None:
            sigma = (np.array(np.asarray(block_size)).astype(float) - 1) / 6.0
        else:
            sigma = param
        thresh_image = _gaussian.gaussian_filter(image, sigma, mode=mode, cval=cval)
    elif method == 'mean':
        size = np.asarray(block_size)
        size = int(size) if size.ndim == 0 else tuple(int(s) for s in size)
        thresh_image = _smooth.uniform_filter(image, size, mode=mode, cval=cval)
    elif method == 'median':
        raise OutOfScopeError("method='median' is a rank filter; it is outside the B200 correlation path")
    else:
        raise ValueError("Invalid method specified. Please use `generic`, `gaussian`, `mean`, or `median`.")
    return thresh_image - offset
