"""CPU arithmetic backend for the slab engine -- TEST INFRASTRUCTURE ONLY.

Plugs the oracle (``oracle/``) into ``dask_image_b200.sharded.ShardedVolume`` so
that the partitioning / halo exchange / boundary logic of the multi-GPU engine
can be exercised on ``gloo`` without a GPU.  It implements the C ABI's slab
contract literally: the resident block (pads + planes) is filtered as if it
were a whole array -- ``mode`` then acts at the ends of the resident range --
and the pads are trimmed off.
"""
import numpy as np
import torch

from oracle import oracle as orc

EPI_STORE, EPI_ACC, EPI_SQ_STORE, EPI_SQ_ACC, EPI_SQ_ACC_SQRT = range(5)


def _np(t):
    # torch.uint16/32/64 <-> numpy through raw bytes
    return t.numpy() if t.dtype not in (torch.uint16, torch.uint32, torch.uint64) else \
        t.view(torch.uint8).numpy().view({torch.uint16: np.uint16, torch.uint32: np.uint32,
                                          torch.uint64: np.uint64}[t.dtype]).reshape(tuple(t.shape))


def _fold(dst, r, op):
    with np.errstate(all="ignore"):
        if op == EPI_STORE:
            dst[...] = r
        elif op == EPI_ACC:
            dst += r
        else:
            np.multiply(r, r, r)
            if op == EPI_SQ_STORE:
                dst[...] = r
            else:
                dst += r
                if op == EPI_SQ_ACC_SQRT:
                    np.sqrt(dst, dst, casting="unsafe")


class OracleBackend:
    name = "oracle"

    @staticmethod
    def _block(src_full, s0, n, pad):
        return np.ascontiguousarray(_np(src_full)[s0 - pad[0]:s0 + n + pad[1]])

    @classmethod
    def correlate1d(cls, src_full, s0, dst_full, d0, n, pad, shape, np_dtype, ps, cval, epilogue, flags=None):
        blk = cls._block(src_full, s0, n, pad if ps.axis == 0 else (0, 0))
        r = orc.correlate1d(blk, ps.weights, ps.axis, ps.mode, cval, ps.origin)
        if ps.axis == 0:
            r = r[pad[0]:pad[0] + n]
        _fold(_np(dst_full)[d0:d0 + n], r.copy(), epilogue)

    @classmethod
    def uniform1d(cls, src_full, s0, dst_full, d0, n, pad, shape, np_dtype, ps, cval, flags=None):
        blk = cls._block(src_full, s0, n, pad if ps.axis == 0 else (0, 0))
        r = orc.uniform_filter1d(blk, ps.size, ps.axis, ps.mode, cval, ps.origin)
        if ps.axis == 0:
            r = r[pad[0]:pad[0] + n]
        _np(dst_full)[d0:d0 + n] = r

    @classmethod
    def correlate_nd(cls, src_full, s0, dst_full, d0, n, pad, shape, np_dtype, w, origins, mode, cval,
                     flags=None):
        blk = cls._block(src_full, s0, n, pad)
        r = orc.correlate(blk, w, mode, cval, origins)
        _np(dst_full)[d0:d0 + n] = r[pad[0]:pad[0] + n]

    @classmethod
    def minmax1d(cls, src_full, s0, dst_full, d0, n, pad, shape, np_dtype, ps, cval, flags=None):
        import scipy.ndimage as ndi

        blk = cls._block(src_full, s0, n, pad if ps.axis == 0 else (0, 0))
        f = ndi.maximum_filter1d if ps.kind == "max" else ndi.minimum_filter1d
        r = f(blk, ps.size, axis=ps.axis, mode=ps.mode, cval=cval, origin=ps.origin)
        if ps.axis == 0:
            r = r[pad[0]:pad[0] + n]
        _np(dst_full)[d0:d0 + n] = r

    @classmethod
    def minmax_nd(cls, src_full, s0, dst_full, d0, n, pad, shape, np_dtype, fp, origins, is_max, mode, cval,
                  flags=None):
        import scipy.ndimage as ndi

        blk = cls._block(src_full, s0, n, pad)
        f = ndi.maximum_filter if is_max else ndi.minimum_filter
        r = f(blk, footprint=fp, mode=mode, cval=cval, origin=origins)
        _np(dst_full)[d0:d0 + n] = r[pad[0]:pad[0] + n]

    @staticmethod
    def combine(dst, src, np_dtype, op):
        _fold(_np(dst), _np(src).copy(), op)
