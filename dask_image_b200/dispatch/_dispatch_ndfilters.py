"""Dispatch tables of the ndfilters hot path and the ``B200Array``
registrations (mirrors dask_image/dispatch/_dispatch_ndfilters.py:28-43 for the
dispatchers SURVEY.md section 8 puts in scope).

Only ``B200Array`` is registered in this package's own tables: a host
(numpy) array raises ``TypeError`` instead of silently running on the CPU.

``register_b200(dispatchers)`` performs the registration a maintainer would add
to the reference; it is applied to this package's own
tables at import and, when ``dask_image`` is importable, to the reference's
tables too, so unmodified ``dask_image.ndfilters.*`` calls route ``B200Array``
chunks to the CUDA kernels.
"""
from .. import ndimage as _b200_ndimage
from .. import sharded as _b200_sharded
from .. import streaming as _b200_streaming
from .._array import B200Array
from ._dispatcher import Dispatcher

_NAMES = (
    "convolve", "correlate", "laplace", "prewitt", "sobel", "gaussian_filter",
    "gaussian_gradient_magnitude", "gaussian_laplace", "uniform_filter",
    "minimum_filter", "maximum_filter",
)

__all__ = ["dispatch_" + n for n in _NAMES] + ["register_b200"]

dispatch_convolve = Dispatcher(name="dispatch_convolve")
dispatch_correlate = Dispatcher(name="dispatch_correlate")
dispatch_laplace = Dispatcher(name="dispatch_laplace")
dispatch_prewitt = Dispatcher(name="dispatch_prewitt")
dispatch_sobel = Dispatcher(name="dispatch_sobel")
dispatch_gaussian_filter = Dispatcher(name="dispatch_gaussian_filter")
dispatch_gaussian_gradient_magnitude = Dispatcher(name="dispatch_gaussian_gradient_magnitude")
dispatch_gaussian_laplace = Dispatcher(name="dispatch_gaussian_laplace")
dispatch_uniform_filter = Dispatcher(name="dispatch_uniform_filter")
dispatch_minimum_filter = Dispatcher(name="dispatch_minimum_filter")
dispatch_maximum_filter = Dispatcher(name="dispatch_maximum_filter")


def _factory(func):
    def b200_factory(*args, **kwargs):
        return func

    b200_factory.__name__ = "b200_" + func.__name__
    return b200_factory


def register_b200(dispatchers):
    """Register the B200 chunk functions for ``B200Array`` in a mapping
    ``{"gaussian_filter": <Dispatcher>, ...}`` (unknown names are ignored)."""
    for name in _NAMES:
        d = dispatchers.get(name)
        if d is not None:
            d.register(B200Array)(_factory(getattr(_b200_ndimage, name)))
            # a whole multi-GPU volume: the slab engine plays map_overlap's part
            d.register(_b200_sharded.ShardedVolume)(_factory(_b200_sharded.FUNCTIONS[name]))
            # a host-resident volume streamed through the GPU (H2D | filter | D2H overlapped)
            d.register(_b200_streaming.HostVolume)(_factory(_b200_streaming.FUNCTIONS[name]))


_OWN = {n: globals()["dispatch_" + n] for n in _NAMES}
register_b200(_OWN)


def register_with_dask_image():
    """Hook ``B200Array`` into an installed dask-image.  Returns True if done."""
    try:
        from dask_image.dispatch import _dispatch_ndfilters as ref
    except ImportError:
        return False
    register_b200({n: getattr(ref, "dispatch_" + n, None) for n in _NAMES})
    return True


register_with_dask_image()
