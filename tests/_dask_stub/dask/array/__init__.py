"""``dask.array`` as far as the reference's ndfilters tests use it: ``from_array``, ``stack``, ``concatenate``,
``utils.assert_eq``.  Arrays are ``dask_image_b200.HaloChunked`` objects -- chunks live in ``B200Array`` buffers, so the
reference's dispatchers route every chunk to this repository's chunk functions."""
import numpy as np
import torch

from dask_image_b200 import B200Array
from dask_image_b200._overlap import HaloChunked

from . import utils  # noqa: F401


class Array(HaloChunked):
    """What ``from_array`` returns: HaloChunked plus the comparison the threshold tests apply to it."""

    def __gt__(self, other):
        return _Host(np.asarray(self) > np.asarray(other))


class _Host:
    def __init__(self, a):
        self._a = a

    def compute(self):
        return self._a


def from_array(a, chunks=None):
    a = np.asarray(a)
    utils.note_data(a)
    if chunks is None:
        chunks = a.shape
    return Array(B200Array.from_numpy(np.ascontiguousarray(a)), chunks)


def _tensors(seq):
    return [x.compute().tensor if hasattr(x, "compute") else B200Array.from_numpy(np.asarray(x)).tensor for x in seq]


def stack(seq, axis=0):
    seq = list(seq)
    out = B200Array(torch.stack(_tensors(seq), dim=axis))
    chunks = list(seq[0].chunksize)
    chunks.insert(axis, 1)
    return HaloChunked(out, tuple(chunks))


def concatenate(seq, axis=0):
    seq = list(seq)
    return HaloChunked(B200Array(torch.cat(_tensors(seq), dim=axis)), seq[0].chunksize)
