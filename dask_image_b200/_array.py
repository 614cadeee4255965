"""``B200Array`` -- the array type registered with ``dask_image.dispatch``.

It is a thin owner of one C-contiguous device buffer (a CUDA ``torch.Tensor``;
PyTorch is used for device-memory ownership and streams only) with numpy
dtype/shape semantics, i.e. the role ``cupy.ndarray`` plays for the
reference's lazily registered GPU backend
(dask_image/dispatch/_dispatch_ndfilters.py:132-139).
"""
import numpy as np
import torch

_NP_TO_TORCH = {
    np.dtype(np.uint8): torch.uint8, np.dtype(np.int8): torch.int8,
    np.dtype(np.uint16): torch.uint16, np.dtype(np.int16): torch.int16,
    np.dtype(np.uint32): torch.uint32, np.dtype(np.int32): torch.int32,
    np.dtype(np.uint64): torch.uint64, np.dtype(np.int64): torch.int64,
    np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64,
}
_TORCH_TO_NP = {v: k for k, v in _NP_TO_TORCH.items()}


def torch_dtype(np_dtype):
    try:
        return _NP_TO_TORCH[np.dtype(np_dtype)]
    except KeyError:
        # same condition scipy reports for float16 / complex / object input
        raise RuntimeError("data type not supported")


def default_device():
    if not torch.cuda.is_available():
        raise RuntimeError("dask_image_b200 needs a CUDA device (sm_100a); none is visible "
                           "and there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


# numpy functions that work on B200Array without leaving the device (the
# __array_function__ protocol, NEP 18): what dask needs to assemble halo'd blocks
# from B200Array chunks (concatenate / stack / *_like / slicing) -- SURVEY.md
# 8(f) rank 3.  Anything else raises TypeError instead of silently copying to
# the host.
_HANDLED = {}


def _implements(np_function):
    def deco(fn):
        _HANDLED[np_function] = fn
        return fn

    return deco


class B200Array:
    """N-D array resident in B200 HBM.

    Parameters
    ----------
    tensor : torch.Tensor
        CUDA tensor; made contiguous if it is not.
    """

    __slots__ = ("tensor",)
    # make numpy defer binary ops / ufuncs to us instead of coercing
    __array_priority__ = 100

    def __init__(self, tensor):
        if not isinstance(tensor, torch.Tensor):
            raise TypeError("B200Array wraps a torch.Tensor; use B200Array.from_numpy for host data")
        if tensor.dtype not in _TORCH_TO_NP:
            raise RuntimeError("data type not supported")
        if tensor.device.type != "cuda":
            raise RuntimeError("B200Array needs a CUDA tensor; there is no CPU fallback")
        self.tensor = tensor if tensor.is_contiguous() else tensor.contiguous()

    # ---- construction / export -------------------------------------------
    @classmethod
    def from_numpy(cls, a, device=None, non_blocking=False):
        a = np.ascontiguousarray(a)
        torch_dtype(a.dtype)
        dev = device if device is not None else default_device()
        t = torch.from_numpy(a) if a.flags.writeable else torch.from_numpy(a.copy())
        return cls(t.to(dev, non_blocking=non_blocking))

    @classmethod
    def empty(cls, shape, dtype, device=None):
        dev = device if device is not None else default_device()
        return cls(torch.empty(tuple(int(s) for s in shape), dtype=torch_dtype(dtype), device=dev))

    def empty_like(self):
        return B200Array(torch.empty_like(self.tensor))

    @classmethod
    def from_device_array(cls, obj):
        """Zero-copy view of another library's device array: anything that speaks DLPack (``__dlpack__``: cupy,
        torch, jax, numba >= 0.55) or the CUDA Array Interface (``__cuda_array_interface__``: cupy, numba, rmm).
        The buffer must be C-contiguous (it is copied on the device otherwise); ``obj`` is kept alive by the view.
        This is how chunks of the reference's cupy backend (dask_image/dispatch/_dispatch_ndfilters.py:132-139) enter
        this path without a round trip through the host."""
        if isinstance(obj, torch.Tensor):
            return cls(obj)
        cai = getattr(obj, "__cuda_array_interface__", None)
        if cai is None and hasattr(obj, "__dlpack__") and hasattr(obj, "__dlpack_device__"):
            if int(obj.__dlpack_device__()[0]) != 2:      # kDLCUDA
                raise TypeError("expected an array in CUDA device memory (use B200Array.from_numpy for host data)")
            return cls(torch.from_dlpack(obj))
        if cai is None:
            raise TypeError("expected a device array exposing __dlpack__ or __cuda_array_interface__")
        dt = np.dtype(cai["typestr"])
        torch_dtype(dt)
        if dt.kind == "u" and dt.itemsize > 1:
            # torch reads the interface for the signed types only: take the bits, view them back
            signed = dict(cai, typestr=np.dtype("i%d" % dt.itemsize).str)
            proxy = type("_CAI", (), {"__cuda_array_interface__": signed, "_owner": obj})()
            return cls(torch.as_tensor(proxy, device="cuda").view(torch_dtype(dt)))
        return cls(torch.as_tensor(obj, device="cuda"))

    def to_numpy(self):
        return self.tensor.cpu().numpy()

    # ---- zero-copy export to other GPU libraries -----------------------------
    def __dlpack__(self, *args, **kwargs):
        return self.tensor.__dlpack__(*args, **kwargs)

    def __dlpack_device__(self):
        return self.tensor.__dlpack_device__()

    @property
    def __cuda_array_interface__(self):
        """CUDA Array Interface v3 (``cupy.asarray(x)`` / ``numba.cuda.as_cuda_array(x)`` view the buffer in place)."""
        if self.tensor.device.type != "cuda":
            raise AttributeError("__cuda_array_interface__")
        return {"shape": self.shape, "typestr": self.dtype.str, "strides": None, "version": 3,
                "data": (self.tensor.data_ptr() if self.size else 0, False)}

    def __array__(self, dtype=None, copy=None):
        a = self.to_numpy()
        return a if dtype is None else a.astype(dtype, copy=False)

    # ---- numpy-like surface ----------------------------------------------
    @property
    def shape(self):
        return tuple(int(s) for s in self.tensor.shape)

    @property
    def ndim(self):
        return self.tensor.dim()

    @property
    def dtype(self):
        return _TORCH_TO_NP[self.tensor.dtype]

    @property
    def size(self):
        return int(self.tensor.numel())

    @property
    def nbytes(self):
        return self.size * self.dtype.itemsize

    @property
    def device(self):
        return self.tensor.device

    @property
    def data_ptr(self):
        return self.tensor.data_ptr()

    def __len__(self):
        if self.ndim == 0:
            raise TypeError("len() of unsized object")
        return self.shape[0]

    def __getitem__(self, key):
        t = self.tensor[key]
        if not isinstance(t, torch.Tensor):
            t = torch.as_tensor(t, device=self.tensor.device)
        return B200Array(t)

    def copy(self):
        return B200Array(self.tensor.clone())

    def astype(self, dtype, copy=True):
        td = torch_dtype(dtype)
        if td == self.tensor.dtype and not copy:
            return self
        return B200Array(self.tensor.to(td, copy=True))

    def reshape(self, *shape):
        if len(shape) == 1 and not np.isscalar(shape[0]):
            shape = tuple(shape[0])
        return B200Array(self.tensor.reshape(shape))

    def __sub__(self, other):
        # threshold_local subtracts a scalar offset from the filtered image
        # (dask_image/ndfilters/_threshold.py:97)
        if isinstance(other, B200Array):
            return B200Array(self.tensor - other.tensor)
        return B200Array(self.tensor - other)

    def __array_function__(self, func, types, args, kwargs):
        handler = _HANDLED.get(func)
        if handler is None or not all(issubclass(t, B200Array) for t in types):
            return NotImplemented
        return handler(*args, **kwargs)

    def transpose(self, *axes):
        if len(axes) == 1 and not np.isscalar(axes[0]):
            axes = tuple(axes[0])
        if not axes:
            axes = tuple(range(self.ndim))[::-1]
        return B200Array(self.tensor.permute(*axes))

    def __repr__(self):
        return "B200Array(shape=%s, dtype=%s, device=%s)" % (self.shape, self.dtype, self.device)


def as_b200(x, like=None):
    """Coerce host data to a ``B200Array`` (device of ``like`` if given)."""
    if isinstance(x, B200Array):
        return x
    if isinstance(x, torch.Tensor):
        return B200Array(x)
    if not isinstance(x, np.ndarray) and (hasattr(x, "__cuda_array_interface__") or (
            hasattr(x, "__dlpack_device__") and x.__dlpack_device__()[0] == 2)):      # 2 = kDLCUDA
        return B200Array.from_device_array(x)
    dev = like.device if like is not None else None
    return B200Array.from_numpy(np.asarray(x), device=dev)


@_implements(np.concatenate)
def _concatenate(arrays, axis=0, out=None, **kwargs):
    if out is not None:
        raise TypeError("concatenate(out=...) is not supported for B200Array")
    return B200Array(torch.cat([as_b200(a).tensor for a in arrays], dim=0 if axis is None else int(axis)))


@_implements(np.stack)
def _stack(arrays, axis=0, out=None, **kwargs):
    if out is not None:
        raise TypeError("stack(out=...) is not supported for B200Array")
    return B200Array(torch.stack([as_b200(a).tensor for a in arrays], dim=int(axis)))


def _like(prototype, dtype, shape):
    dt = prototype.tensor.dtype if dtype is None else torch_dtype(dtype)
    shp = prototype.shape if shape is None else tuple(np.atleast_1d(shape).tolist())
    return dt, shp


@_implements(np.empty_like)
def _empty_like(prototype, dtype=None, order="K", subok=True, shape=None, **kwargs):
    dt, shp = _like(prototype, dtype, shape)
    return B200Array(torch.empty(shp, dtype=dt, device=prototype.device))


@_implements(np.zeros_like)
def _zeros_like(a, dtype=None, order="K", subok=True, shape=None, **kwargs):
    dt, shp = _like(a, dtype, shape)
    return B200Array(torch.zeros(shp, dtype=dt, device=a.device))


@_implements(np.shape)
def _shape(a):
    return a.shape


@_implements(np.ndim)
def _ndim(a):
    return a.ndim


@_implements(np.transpose)
def _transpose(a, axes=None):
    return a.transpose() if axes is None else a.transpose(axes)

