"""``dask.array.utils.assert_eq``: same shape, same dtype, same values.  Stricter than dask's (which compares with
``allclose``) whenever the chunks are host tensors: below the CPU test harness the arithmetic is the oracle and the
results must be bit-identical to scipy.  On a device the float kernels are held to the repository's tolerance
(tests/_helpers.py): ``rtol * |ref| + rtol / 5 * scale``, the scale being the largest magnitude among the two arrays
AND the inputs handed to ``from_array`` in the running test -- a derivative kernel applied to a ramp is exactly 0 in
scipy (it folds x[-j] - x[j] first) while a direct-form kernel leaves rounding noise relative to the DATA."""
import os

import numpy as np

FORCE_TOLERANCE = bool(os.environ.get("B200_STUB_FORCE_TOLERANCE"))      # exercise the device branch on the CPU
DATA_SCALE = [0.0]      # reset per test by b200_refsuite_plugin, raised by from_array


def note_data(a):
    if a.size and a.dtype.kind in "fiu":
        DATA_SCALE[0] = max(DATA_SCALE[0], float(np.abs(a.astype(np.float64)).max()))


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
    if (dev_a or dev_b or FORCE_TOLERANCE) and a.dtype.kind == "f":
        rtol = 1e-5 if a.dtype == np.float32 else 1e-12
        scale = max(float(np.abs(a).max(initial=0.0)), float(np.abs(b).max(initial=0.0)), DATA_SCALE[0])
        err = np.abs(a.astype(np.float64) - b.astype(np.float64))
        bound = rtol * np.minimum(np.abs(a), np.abs(b)).astype(np.float64) + rtol / 5 * scale
        assert np.all((err <= bound) | (np.isnan(a) & np.isnan(b))), "max |a - b| = %g (scale %g)" % (err.max(), scale)
    else:
        np.testing.assert_array_equal(a, b)
    return True
