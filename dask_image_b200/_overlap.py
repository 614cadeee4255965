"""``HaloChunked`` -- a minimal eager stand-in for a chunked dask array whose
chunks are :class:`B200Array` objects.

dask is not installable in the build image, so the parity tests cannot wrap
inputs with ``dask.array.from_array(a, chunks=...)`` the way the reference's
tests do.  This class exposes exactly the surface the ndfilters wrappers touch
(``ndim``, ``shape``, ``dtype``, ``_meta``, ``map_overlap``) and implements
``map_overlap(depth, boundary in {"none", "periodic"})`` with dask's semantics:
every chunk is extended by ``depth`` voxels taken from its neighbours (nothing
is added past the outer faces for ``"none"``; the opposite end's planes for
``"periodic"``), the chunk function runs on the halo'd block and ``depth`` is
trimmed off again.  It drives the per-chunk drop-in boundary (the functions in
``dask_image_b200.ndimage``) with the same kwargs dask would pass.
"""
import itertools
import numbers

import torch

from ._array import B200Array, as_b200


class HaloChunked:
    def __init__(self, array, chunks):
        self._array = as_b200(array)
        nd = self._array.ndim
        if isinstance(chunks, int):
            chunks = (chunks,) * nd
        if len(chunks) != nd:
            raise ValueError("chunks must have one entry per dimension")
        self.chunksize = tuple(int(c) for c in chunks)
        self._meta = B200Array(torch.empty((0,) * nd, dtype=self._array.tensor.dtype,
                                           device=self._array.device))

    # ---- the surface the wrappers use ------------------------------------
    @property
    def ndim(self):
        return self._array.ndim

    @property
    def shape(self):
        return self._array.shape

    @property
    def dtype(self):
        return self._array.dtype

    def __len__(self):
        return len(self._array)

    def compute(self):
        return self._array

    def __array__(self, dtype=None, copy=None):
        return self._array.__array__(dtype)

    def __getitem__(self, key):
        sub = self._array[key]
        if not sub.ndim:
            return sub
        keys = key if isinstance(key, tuple) else (key,)
        if all(isinstance(k, numbers.Integral) for k in keys):
            chunks = self.chunksize[len(keys):]             # d[i]: the leading axes are dropped
        elif key is None:
            chunks = (1,) + self.chunksize                  # d[None]: a new leading axis
        else:
            chunks = sub.shape                              # anything else: one chunk
        return HaloChunked(sub, chunks)

    def astype(self, dtype):
        return HaloChunked(self._array.astype(dtype), self.chunksize)

    def __sub__(self, other):
        return HaloChunked(self._array - other, self.chunksize)

    def map_overlap(self, func, depth, boundary, dtype=None, meta=None, **kwargs):
        nd = self.ndim
        # dask accepts a scalar or a per-axis mapping for both arguments
        depth = [int(depth[d]) if hasattr(depth, "__getitem__") else int(depth) for d in range(nd)]
        bnd = [boundary if (boundary is None or isinstance(boundary, str)) else boundary[d]
               for d in range(nd)]
        src = self._array.tensor
        shift = [0] * nd
        for d in range(nd):
            if bnd[d] == "periodic" and depth[d] > 0:
                n = src.shape[d]
                idx = torch.arange(-depth[d], n + depth[d], device=src.device) % n
                src = src.index_select(d, idx)
                shift[d] = depth[d]
            elif bnd[d] not in ("none", None, "periodic"):
                raise NotImplementedError("boundary=%r" % (bnd[d],))
        out = torch.empty_like(self._array.tensor)
        spans = [[(s, min(s + self.chunksize[d], self.shape[d]))
                  for s in range(0, self.shape[d], self.chunksize[d])] for d in range(nd)]
        for job in itertools.product(*spans):
            src_sl, trim_sl, dst_sl = [], [], []
            for d, (s, e) in enumerate(job):
                lo = max(s + shift[d] - depth[d], 0)
                hi = min(e + shift[d] + depth[d], src.shape[d])
                src_sl.append(slice(lo, hi))
                t0 = s + shift[d] - lo
                trim_sl.append(slice(t0, t0 + (e - s)))
                dst_sl.append(slice(s, e))
            block = B200Array(src[tuple(src_sl)].contiguous())
            res = func(block, **kwargs)
            if type(res) is not B200Array or res.shape != block.shape or res.dtype != block.dtype:
                raise TypeError("chunk function must return a B200Array of the chunk's shape/dtype")
            out[tuple(dst_sl)] = res.tensor[tuple(trim_sl)]
        return HaloChunked(B200Array(out), self.chunksize)


def from_array(a, chunks):
    """``dask.array.from_array`` look-alike returning a :class:`HaloChunked`."""
    return HaloChunked(a, chunks)
