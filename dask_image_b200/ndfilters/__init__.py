"""Drop-in for the correlation family of ``dask_image.ndfilters``: same names,
signatures, validation and exception classes (dask_image/ndfilters/__init__.py).

The wrapper modules import ``scipy.ndimage`` for ONE purpose, the one the
reference has (dask_image/ndfilters/_utils.py:10-48): every public function
takes its ``__name__`` and docstring from the scipy function it mirrors.  No
scipy function is ever called on this path -- all arithmetic runs in the CUDA
library, and there is no CPU fallback."""
from ._conv import convolve, correlate
from ._diff import laplace
from ._edge import prewitt, sobel
from ._gaussian import (gaussian, gaussian_filter, gaussian_gradient_magnitude,
                        gaussian_laplace)
from ._order import maximum_filter, minimum_filter
from ._smooth import uniform_filter
from ._threshold import threshold_local
from ._unsupported import (OutOfScopeError, generic_filter, median_filter, percentile_filter,
                           rank_filter)

__all__ = [
    "convolve", "correlate", "laplace", "prewitt", "sobel", "gaussian", "gaussian_filter",
    "gaussian_gradient_magnitude", "gaussian_laplace", "uniform_filter", "threshold_local",
    "minimum_filter", "maximum_filter",
    # present under the reference's names, raising OutOfScopeError (rank statistics / Python callables: SURVEY 8)
    "generic_filter", "median_filter", "rank_filter", "percentile_filter", "OutOfScopeError",
]

for _f in (convolve, correlate, laplace, prewitt, sobel, gaussian, gaussian_filter,
           gaussian_gradient_magnitude, gaussian_laplace, uniform_filter, threshold_local,
           minimum_filter, maximum_filter, generic_filter, median_filter, rank_filter, percentile_filter):
    _f.__module__ = __name__
del _f
