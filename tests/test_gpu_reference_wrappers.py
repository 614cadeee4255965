"""The UNMODIFIED reference wrappers driving the CUDA kernels.

``dask_image.ndfilters`` is imported as it is -- from ``baseline/_ref``, the pip-installed copy of the reference that
``__graft_entry__.build()`` makes in the build container and that travels to the GPU box with the snapshot -- with a
stand-in for the one dask class that path imports (``dask.utils.Dispatch``, tests/_reference_stub.py).
``register_with_dask_image()`` adds the ``B200Array`` registrations to the REFERENCE's dispatch tables; the reference's
public functions are then called on ``HaloChunked`` arrays of device chunks (the eager ``map_overlap`` stand-in):
its validation, halo depths, dispatch and per-chunk kwargs run unchanged, the chunk function behind them is this
package's, the arithmetic is the sm_100a kernels through the C ABI.  Values are checked against whole-array scipy
(integers bit for bit).  Mirrors ref:tests/test_dask_image/test_ndfilters/test_cupy_ndfilters.py:13-103.
Skips when no copy of the reference is present, or when a real dask is (then tests/test_gpu_dask.py applies).
"""
import numpy as np
import pytest
import scipy.ndimage as ndi

import _reference_stub
from _helpers import assert_parity

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(_reference_stub.unavailable_reason() is not None,
                                 reason=str(_reference_stub.unavailable_reason()))]


@pytest.fixture
def ref_ndfilters():
    import dask_image_b200

    with _reference_stub.reference_ndfilters() as ref:
        yield ref, dask_image_b200


@pytest.mark.parametrize("dtype", [np.float32, np.uint16])
@pytest.mark.parametrize("chunks", [(20, 16), (13, 48)])
def test_reference_wrappers_drive_the_kernels(ref_ndfilters, dtype, chunks):
    from dask_image_b200 import _lib

    ref, b200 = ref_ndfilters
    rng = np.random.default_rng(12)
    a = (rng.random((40, 48)) * 1000).astype(dtype)
    exact = np.dtype(dtype).kind != "f"
    d = b200.from_array(a, chunks=chunks)
    w = rng.random((3, 5)) - 0.3
    wd = b200.B200Array.from_numpy(w)             # ref:dask_image/dispatch/_utils.py:23-25: same array type
    calls = [
        ("gaussian_filter", (1.0,), {}),
        ("gaussian", ((1.5, 0.7),), dict(order=(0, 1), mode="mirror", truncate=3.0)),
        ("gaussian_gradient_magnitude", (1.0,), {}),
        ("gaussian_laplace", (1.2,), dict(mode="nearest")),
        ("uniform_filter", (3,), {}),
        ("uniform_filter", ((4, 7),), dict(origin=(1, -2), mode="constant", cval=5.0)),
        ("laplace", (), {}),
        ("prewitt", (), {}),
        ("sobel", (), dict(axis=0)),
        ("convolve", (wd,), {}),
        ("correlate", (wd,), dict(mode="wrap", origin=(1, -1))),   # periodic halo + constant per chunk
        ("minimum_filter", (), dict(size=3)),
        ("maximum_filter", (), dict(footprint=np.array([[0, 1, 0], [1, 1, 1], [0, 1, 0]]))),
    ]
    for name, args, kw in calls:
        _lib.launch_count(reset=True)
        got = np.asarray(getattr(ref, name)(d, *args, **kw).compute())
        assert _lib.launch_count() > 0, name            # the CUDA path ran, not scipy
        sargs = tuple(w if x is wd else x for x in args)
        want = getattr(ndi, "gaussian_filter" if name == "gaussian" else name)(a, *sargs, **kw)
        assert_parity(got, want, exact=exact or "filter" in name and name.startswith(("min", "max")), what=name,
                      data=a)
    with pytest.raises(ValueError, match="Array types must be compatible"):
        ref.convolve(d, w)
    # out of scope: no registration for B200Array -> the reference's dispatcher says so
    with pytest.raises(TypeError):
        ref.median_filter(d, size=3)


def test_reference_threshold_local_on_device(ref_ndfilters):
    ref, b200 = ref_ndfilters
    a = (np.random.default_rng(4).random((30, 32)) * 255).astype(np.uint8)
    d = b200.from_array(a, chunks=(15, 16))
    got = np.asarray(ref.threshold_local(d, 7, offset=2.5, mode="mirror").compute())
    want = ndi.gaussian_filter(a.astype(np.float64), (7 - 1) / 6.0, mode="mirror") - 2.5
    assert_parity(got, want, data=a)
