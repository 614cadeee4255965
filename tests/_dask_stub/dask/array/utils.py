"""``dask.array.utils.assert_eq``: same shape, same dtype, same values.  Stricter than dask's (which compares with
``allclose``) whenever the chunks are host tensors: below the CPU test harness the arithmetic is the oracle and the
results must be bit-identical to scipy; on a device the float kernels are held to the repository's tolerance."""
import numpy as np


def _host(x):
    on_device = False
    if hasattr(x, "compute"):
        x = x.compute()
    if hasattr(x, "tensor"):
        on_device = x.tensor.is_cuda
        x = x.to_numpy()
    return np.asarray(x), on_device


def assert_eq(a, b, **kwargs):
    a, dev_a = _host(a)
    b, dev_b = _host(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    assert a.dtype == b.dtype, (a.dtype, b.dtype)
    if (dev_a or dev_b) and a.dtype.kind == "f":
        rtol = 1e-5 if a.dtype == np.float32 else 1e-12
        scale = max(float(np.abs(a).max(initial=0.0)), float(np.abs(b).max(initial=0.0)))
        np.testing.assert_allclose(a, b, rtol=rtol, atol=rtol / 5 * scale, equal_nan=True)
    else:
        np.testing.assert_array_equal(a, b)
    return True
