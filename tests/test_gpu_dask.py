"""The reference itself driving the kernels: unmodified ``dask_image.ndfilters.*`` on a dask array whose
chunks are ``B200Array`` objects.  Mirrors the reference's cupy smoke tests
(ref:tests/test_dask_image/test_ndfilters/test_cupy_ndfilters.py:13-103) but asserts VALUES against scipy.

dask and dask-image are not installed in the build image nor on the GPU boxes of this pool (probed in round 2:
``import dask`` -> ModuleNotFoundError, gpurun_out/r2_dask_probe.log; there is no network to install them), so
this module skips there.  It runs wherever both are importable.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

dask = pytest.importorskip("dask")
da = pytest.importorskip("dask.array")
dask_image_ndfilters = pytest.importorskip("dask_image.ndfilters")


@pytest.fixture(scope="module")
def b200():
    import dask_image_b200
    from dask_image_b200.dispatch._dispatch_ndfilters import register_with_dask_image

    assert register_with_dask_image()
    try:   # let dask treat B200Array as a chunk type (duck array protocol)
        from dask.array.chunk_types import register_chunk_type

        register_chunk_type(dask_image_b200.B200Array)
    except Exception:
        pass
    return dask_image_b200


def _device_dask_array(b200, a, chunks):
    d = da.from_array(a, chunks=chunks)
    return d.map_blocks(b200.B200Array.from_numpy, dtype=a.dtype,
                        meta=b200.B200Array.from_numpy(np.empty((0,) * a.ndim, dtype=a.dtype)))


def _host(result):
    out = result.compute()
    return out.to_numpy() if hasattr(out, "to_numpy") else np.asarray(out)


@pytest.mark.parametrize("name, kwargs", [
    ("gaussian_filter", dict(sigma=1.0)),
    ("gaussian", dict(sigma=1.0)),
    ("gaussian_gradient_magnitude", dict(sigma=1.0)),
    ("gaussian_laplace", dict(sigma=1.0)),
    ("uniform_filter", dict(size=3)),
    ("laplace", dict()),
    ("prewitt", dict()),
    ("sobel", dict()),
])
def test_reference_wrappers_drive_the_kernels(b200, name, kwargs):
    import scipy.ndimage as ndi

    from dask_image_b200 import _lib

    a = np.arange(40 * 48, dtype=np.float32).reshape(40, 48) / 7.0
    d = _device_dask_array(b200, a, chunks=(20, 16))
    _lib.launch_count(reset=True)
    got = _host(getattr(dask_image_ndfilters, name)(d, **kwargs))
    assert _lib.launch_count() > 0                      # the CUDA path ran, not scipy
    ref = getattr(ndi, "gaussian_filter" if name == "gaussian" else name)(a, **kwargs)
    np.testing.assert_allclose(got, ref, rtol=1e-5, atol=1e-5 * np.abs(ref).max())


@pytest.mark.parametrize("name", ["convolve", "correlate"])
def test_reference_conv_wrap_is_periodic(b200, name):
    import scipy.ndimage as ndi

    rng = np.random.default_rng(3)
    a = rng.random((36, 40)).astype(np.float32)
    w = rng.random((3, 5))
    d = _device_dask_array(b200, a, chunks=(12, 20))
    got = _host(getattr(dask_image_ndfilters, name)(d, b200.B200Array.from_numpy(w), mode="wrap"))
    ref = getattr(ndi, name)(a, w, mode="wrap")
    np.testing.assert_allclose(got, ref, rtol=1e-5, atol=1e-5 * np.abs(ref).max())
