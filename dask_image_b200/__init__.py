"""dask_image_b200 -- B200-native drop-in for the ndfilters correlation family
of dask-image (gaussian_filter & derivative variants, uniform_filter,
correlate1d, N-D convolve/correlate): hand-written sm_100a CUDA kernels behind
a C ABI, a ``B200Array`` chunk type registered through the dispatch mechanism,
and an NCCL slab engine replacing ``map_overlap`` for multi-GPU volumes."""
from ._array import B200Array, as_b200  # noqa: F401
from . import ndimage, ndfilters, dispatch  # noqa: F401
from ._overlap import HaloChunked, from_array  # noqa: F401
from .sharded import ShardedVolume  # noqa: F401
from .streaming import HostVolume  # noqa: F401

__version__ = "0.1.0"
