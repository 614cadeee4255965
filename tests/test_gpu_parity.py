"""GPU parity tests (``-m gpu``): the CUDA path, called through the C ABI by the
host mirror of the reference interface, against the CPU oracle, the committed
golden vectors and scipy (the arithmetic the reference runs per chunk).

Parametrisations follow the reference's own tests
(tests/test_dask_image/test_ndfilters/test__gaussian.py, test__smooth.py,
test__conv.py) with ``HaloChunked`` standing in for ``dask.array.from_array``.
"""
import numpy as np
import pytest
import scipy.ndimage as ndi

from _helpers import assert_parity, golden_case, golden_cases
from oracle import oracle as orc

pytestmark = pytest.mark.gpu

CASES = golden_cases()


@pytest.fixture(scope="module")
def b200():
    import dask_image_b200

    return dask_image_b200


def dev(b200, a):
    return b200.B200Array.from_numpy(np.asarray(a))


def host_kwargs(b200, kwargs):
    return dict(kwargs)


# ---------------------------------------------------------------------------
# golden vectors (all dtypes, all modes, origins, even sizes, long radii)
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("i", range(len(CASES)), ids=[c["name"] for c in CASES])
def test_golden(b200, i):
    func, a, kwargs, expected = golden_case(i)
    got = getattr(b200.ndimage, func)(dev(b200, a), **kwargs)
    assert type(got) is b200.B200Array
    assert_parity(got.to_numpy(), expected, what=CASES[i]["name"], data=a)


@pytest.mark.parametrize("i", range(0, len(CASES), 3), ids=[c["name"] for c in CASES[::3]])
def test_golden_exact_mode_is_bitwise(b200, i):
    """B200_FLAG_EXACT: fp64 accumulation in scipy's order -> bit-identical for floats too."""
    from dask_image_b200 import _lib

    func, a, kwargs, expected = golden_case(i)
    old = b200.ndimage.set_default_flags(_lib.FLAG_EXACT)
    try:
        got = getattr(b200.ndimage, func)(dev(b200, a), **kwargs)
    finally:
        b200.ndimage.set_default_flags(old)
    if func in ("uniform_filter", "uniform_filter1d") and a.dtype.kind == "f":
        # scipy carries a running sum along each line; the kernel sums each
        # window directly -- same value up to fp64 rounding, not bitwise
        assert_parity(got.to_numpy(), expected, what=CASES[i]["name"])
    else:
        assert_parity(got.to_numpy(), expected, exact=True, what=CASES[i]["name"])


def test_scipy_docstring_known_answers(b200):
    nd = b200.ndimage
    assert nd.correlate1d(dev(b200, np.array([2, 8, 0, 4, 1, 9, 9, 0])), [1, 3]).to_numpy().tolist() == \
        [8, 26, 8, 12, 7, 28, 36, 9]
    assert nd.convolve1d(dev(b200, np.array([2, 8, 0, 4, 1, 9, 9, 0])), [1, 3]).to_numpy().tolist() == \
        [14, 24, 4, 13, 12, 36, 27, 0]
    assert nd.uniform_filter1d(dev(b200, np.array([2, 8, 0, 4, 1, 9, 9, 0])), size=3).to_numpy().tolist() == \
        [4, 3, 4, 1, 4, 6, 6, 3]
    np.testing.assert_allclose(
        nd.gaussian_filter1d(dev(b200, np.array([1.0, 2.0, 3.0, 4.0, 5.0])), 1).to_numpy(),
        [1.42704095, 2.06782203, 3.0, 3.93217797, 4.57295905], atol=1e-8)
    # scipy/ndimage/_filters.py:824-829 (integer gaussian_filter)
    a = np.arange(50, step=2).reshape((5, 5))
    assert_parity(nd.gaussian_filter(dev(b200, a), sigma=1).to_numpy(), ndi.gaussian_filter(a, sigma=1))


# ---------------------------------------------------------------------------
# the reference's tests, through the public wrappers, chunked like dask would
# ---------------------------------------------------------------------------
GAUSS_FUNCS = [("gaussian_filter", "gaussian"), ("gaussian_filter", "gaussian_filter"),
               ("gaussian_gradient_magnitude", "gaussian_gradient_magnitude"),
               ("gaussian_laplace", "gaussian_laplace")]


@pytest.mark.parametrize("sigma, truncate", [(1.0, 2.0), (1.0, 4.0), (2.0, 2.0), (2.0, 4.0), ((1.0, 2.0), 4.0)])
@pytest.mark.parametrize("sp_name, da_name", GAUSS_FUNCS)
@pytest.mark.parametrize("chunks", [None, (50, 55)])
def test_gaussian_filters_compare(b200, sp_name, da_name, sigma, truncate, chunks):
    # reference: test__gaussian.py:125-154
    s = (100, 110)
    a = np.arange(float(np.prod(s))).reshape(s)
    d = dev(b200, a) if chunks is None else b200.from_array(a, chunks=chunks)
    got = getattr(b200.ndfilters, da_name)(d, sigma, truncate=truncate)
    ref = getattr(orc, sp_name)(a, sigma, truncate=truncate)
    assert all(type(x) is int for x in got.shape)
    assert_parity(np.asarray(got), ref)
    assert_parity(np.asarray(got), getattr(ndi, sp_name)(a, sigma, truncate=truncate))


@pytest.mark.parametrize("sigma, truncate", [(0.0, 0.0), (1.0, 0.0), (0.0, 1.0), (1.0, 2.0), (1.0, 4.0),
                                             (2.0, 2.0), (2.0, 4.0), ((1.0, 2.0), 4.0)])
@pytest.mark.parametrize("order", [0, 1, 2, 3, (0, 1), (2, 3)])
@pytest.mark.parametrize("da_name", ["gaussian", "gaussian_filter"])
def test_gaussian_derivative_filters_compare(b200, da_name, order, sigma, truncate):
    # reference: test__gaussian.py:157-196
    s = (100, 110)
    a = np.arange(float(np.prod(s))).reshape(s)
    d = b200.from_array(a, chunks=(50, 55))
    got = getattr(b200.ndfilters, da_name)(d, sigma, order, truncate=truncate)
    # (a derivative of a linear ramp is EXACTLY zero in scipy's folded sum; the tiled kernels -- which the halo'd
    # chunks, 63 = 55 + 8 columns wide, now take through the plain-load staging -- leave rounding noise relative
    # to the data: the scale of the comparison is the input's)
    assert_parity(np.asarray(got), orc.gaussian_filter(a, sigma, order, truncate=truncate), data=a)


@pytest.mark.parametrize("sigma, truncate", [(0.0, 0.0), (0.0, 1.0), (0.0, 4.0), (1.0, 0.0)])
@pytest.mark.parametrize("order", [0, 1, 2, 3])
def test_gaussian_filters_identity(b200, order, sigma, truncate):
    # reference: test__gaussian.py:38-75
    a = np.arange(140.0).reshape(10, 14)
    d = b200.from_array(a, chunks=(5, 7))
    if order % 2 == 1 and sigma != 0 and truncate == 0:
        pytest.skip("SciPy zeros the result for odd derivatives with truncate == 0 (scipy#7364)")
    got = b200.ndfilters.gaussian_filter(d, sigma, order, truncate=truncate)
    assert_parity(np.asarray(got), a)
    assert_parity(np.asarray(got), ndi.gaussian_filter(a, sigma, order, truncate=truncate))


@pytest.mark.parametrize("da_name", ["gaussian", "gaussian_filter", "gaussian_gradient_magnitude",
                                     "gaussian_laplace", "uniform_filter"])
def test_filter_comprehensions(b200, da_name):
    # reference: test__gaussian.py:101-122, test__smooth.py:45-57
    np.random.seed(0)
    a = np.random.random((3, 12, 14))
    d = b200.from_array(a, chunks=(3, 6, 7))
    f = getattr(b200.ndfilters, da_name)
    args = (1,) if da_name == "uniform_filter" else (1.0,)
    sp = getattr(ndi, "gaussian_filter" if da_name == "gaussian" else da_name)
    for i in range(len(a)):
        assert_parity(np.asarray(f(d[i], *args)), sp(a[i], *args))


@pytest.mark.parametrize("size, origin", [(1, 0), (2, 0), (3, 0), (3, 1), (3, (1, 0)), ((1, 2), 0),
                                          ((3, 2), (1, 0))])
@pytest.mark.parametrize("chunks", [None, (50, 55)])
def test_uniform_compare(b200, size, origin, chunks):
    # reference: test__smooth.py:60-99
    s = (100, 110)
    a = np.arange(float(np.prod(s))).reshape(s)
    d = dev(b200, a) if chunks is None else b200.from_array(a, chunks=chunks)
    got = b200.ndfilters.uniform_filter(d, size, origin=origin)
    assert_parity(np.asarray(got), orc.uniform_filter(a, size, origin=origin))
    assert_parity(np.asarray(got), ndi.uniform_filter(a, size, origin=origin))


DISC = (np.mgrid[-2:3, -2:3] ** 2).sum(axis=0) < 2.5 ** 2
CONV_W = [(np.ones((1, 1)), 0), (np.ones((2, 2)), 0), (np.ones((2, 3)), 0), (np.ones((2, 3)), (0, 1)),
          (np.ones((2, 3)), (0, -1)), (DISC, 0), (DISC, (1, 2)), (DISC, (-1, -2)), (np.ones((5, 5)), 0),
          (np.ones((7, 7)), 0), (np.ones((8, 8)), 0), (np.ones((10, 10)), 0), (np.ones((5, 5)), 2),
          (np.ones((5, 5)), -2)]


@pytest.mark.parametrize("name", ["convolve", "correlate"])
@pytest.mark.parametrize("wi", range(len(CONV_W)))
@pytest.mark.parametrize("chunks", [None, (5, 7)])
def test_convolutions_compare(b200, name, wi, chunks):
    # reference: test__conv.py:88-156 (weights are a B200Array: array types must match)
    weights, origin = CONV_W[wi]
    a = np.arange(140.0).reshape(10, 14)
    d = dev(b200, a) if chunks is None else b200.from_array(a, chunks=chunks)
    got = getattr(b200.ndfilters, name)(d, dev(b200, weights.astype(float)), origin=origin)
    assert_parity(np.asarray(got), getattr(orc, name)(a, weights, origin=origin))


@pytest.mark.parametrize("name", ["convolve", "correlate"])
@pytest.mark.parametrize("wshape", [(1, 5), (5, 1)])
@pytest.mark.parametrize("mode", ["reflect", "wrap", "nearest", "constant", "mirror"])
@pytest.mark.parametrize("chunks", [None, (5, 7)])
def test_convolutions_modes(b200, name, wshape, mode, chunks):
    # reference: test__conv.py:159-191 -- INTEGER array, all five modes; the
    # chunked run goes through the wrap -> periodic rewrite (_conv.py:23-25)
    a = np.arange(140).reshape(10, 14)
    d = dev(b200, a) if chunks is None else b200.from_array(a, chunks=chunks)
    if chunks is not None and mode != "wrap":
        # the reference's own per-chunk semantics for non-wrap modes equal whole-array scipy
        pass
    got = getattr(b200.ndfilters, name)(d, dev(b200, np.ones(wshape)), mode=mode)
    assert_parity(np.asarray(got), getattr(orc, name)(a, np.ones(wshape), mode=mode))
    assert_parity(np.asarray(got), getattr(ndi, name)(a, np.ones(wshape), mode=mode))


@pytest.mark.parametrize("name", ["convolve", "correlate"])
@pytest.mark.parametrize("err_type, weights, origin", [
    (ValueError, np.ones((1,)), 0), (ValueError, np.ones((1, 0)), 0), (RuntimeError, np.ones((1, 1)), (0,)),
    (RuntimeError, np.ones((1, 1)), [(0,)]), (ValueError, np.ones((1, 1)), 1), (TypeError, np.ones((1, 1)), 0.0),
    (TypeError, np.ones((1, 1)), (0.0, 0.0)), (TypeError, np.ones((1, 1)), 1 + 0j),
    (TypeError, np.ones((1, 1)), (0 + 0j, 1 + 0j)),
])
def test_convolutions_params(b200, name, err_type, weights, origin):
    # reference: test__conv.py:12-43
    d = b200.from_array(np.arange(140.0).reshape(10, 14), chunks=(5, 7))
    with pytest.raises(err_type):
        getattr(b200.ndfilters, name)(d, dev(b200, weights), origin=origin)


def test_numpy_input_has_no_cpu_fallback(b200):
    with pytest.raises((TypeError, AttributeError)):
        b200.ndfilters.gaussian_filter(np.zeros((4, 4)), 1.0)


# ---------------------------------------------------------------------------
# edge cases the reference never exercises
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(0, 5), (5, 0), (1, 1), (1, 7), (7, 1), (2, 3, 1, 4)])
def test_empty_and_degenerate_shapes(b200, shape):
    a = np.random.default_rng(0).random(shape)
    nd = b200.ndimage
    assert_parity(nd.gaussian_filter(dev(b200, a), 1.5).to_numpy(), ndi.gaussian_filter(a, 1.5))
    assert_parity(nd.uniform_filter(dev(b200, a), 3).to_numpy(), ndi.uniform_filter(a, 3))
    if a.size:
        w = np.ones((3,) * a.ndim)
        assert_parity(nd.correlate(dev(b200, a), w).to_numpy(), ndi.correlate(a, w))


@pytest.mark.parametrize("mode", ["reflect", "wrap", "nearest", "constant", "mirror"])
def test_radius_much_larger_than_axis(b200, mode):
    a = np.random.default_rng(1).random((3, 5))
    got = b200.ndimage.gaussian_filter(dev(b200, a), 6.0, mode=mode, cval=0.25)
    assert_parity(got.to_numpy(), ndi.gaussian_filter(a, 6.0, mode=mode, cval=0.25))
    got = b200.ndimage.uniform_filter(dev(b200, a), 17, mode=mode, cval=0.25)
    assert_parity(got.to_numpy(), ndi.uniform_filter(a, 17, mode=mode, cval=0.25))


def test_unsigned_wrap_and_truncation(b200):
    # truncation toward zero, negative -> unsigned wraps modulo 2^bits
    a = np.random.default_rng(2).integers(0, 65536, (40, 50), dtype=np.uint16)
    for order in (1, 2):
        got = b200.ndimage.gaussian_filter(dev(b200, a), 2.0, order=order)
        assert_parity(got.to_numpy(), ndi.gaussian_filter(a, 2.0, order=order))
    w = np.array([[-1.5, 2.0], [0.25, -3.0]])
    assert_parity(b200.ndimage.correlate(dev(b200, a), w).to_numpy(), ndi.correlate(a, w))
    a8 = a.astype(np.int8)
    assert_parity(b200.ndimage.correlate(dev(b200, a8), w * 40).to_numpy(), ndi.correlate(a8, w * 40))


def test_grid_mode_aliases(b200):
    a = np.random.default_rng(3).random((9, 11))
    for alias, mode in [("grid-wrap", "wrap"), ("grid-constant", "constant"), ("grid-mirror", "reflect")]:
        got = b200.ndimage.gaussian_filter(dev(b200, a), 2.0, mode=alias, cval=1.0)
        assert_parity(got.to_numpy(), ndi.gaussian_filter(a, 2.0, mode=mode, cval=1.0))
    with pytest.raises(RuntimeError, match="boundary mode not supported"):
        b200.ndimage.gaussian_filter(dev(b200, a), 2.0, mode="bogus")


def test_input_is_not_mutated_and_output_is_new(b200):
    a = np.random.default_rng(4).random((33, 47)).astype(np.float32)
    d = dev(b200, a)
    out = b200.ndimage.gaussian_filter(d, 1.0)
    assert out.tensor.data_ptr() != d.tensor.data_ptr()
    assert np.array_equal(d.to_numpy(), a)


def test_per_axis_modes_and_axes(b200):
    a = np.random.default_rng(5).random((20, 30))
    got = b200.ndimage.gaussian_filter(dev(b200, a), (1.0, 2.0), mode=("wrap", "mirror"))
    assert_parity(got.to_numpy(), ndi.gaussian_filter(a, (1.0, 2.0), mode=("wrap", "mirror")))
    got = b200.ndimage.uniform_filter(dev(b200, a), 5, axes=(1,))
    assert_parity(got.to_numpy(), ndi.uniform_filter(a, 5, axes=(1,)))


class _CaiProducer:
    """A device array that speaks only the CUDA Array Interface (as numba / rmm buffers do)."""

    def __init__(self, tensor, np_dtype):
        self._t, self._dt = tensor, np.dtype(np_dtype)

    @property
    def __cuda_array_interface__(self):
        return {"shape": tuple(self._t.shape), "typestr": self._dt.str, "strides": None, "version": 3,
                "data": (self._t.data_ptr(), False)}


class _DlpackProducer:
    """A device array that speaks only DLPack."""

    def __init__(self, tensor):
        self._t = tensor

    def __dlpack__(self, *args, **kwargs):
        return self._t.__dlpack__(*args, **kwargs)

    def __dlpack_device__(self):
        return self._t.__dlpack_device__()


@pytest.mark.parametrize("dt", ["float32", "uint16", "uint8", "int16", "float64"])
def test_zero_copy_interop_with_other_gpu_libraries(b200, dt):
    """Chunks of another GPU backend (the reference's is cupy, ref:_dispatch_ndfilters.py:132-139) enter and leave
    without a host round trip: DLPack and the CUDA Array Interface, both directions, same device pointer."""
    import torch

    a = (np.random.default_rng(3).random((12, 40)) * 200).astype(dt)
    d = dev(b200, a)
    ptr = d.tensor.data_ptr()
    # export: DLPack and CAI consumers see the same buffer
    assert torch.from_dlpack(d).data_ptr() == ptr
    cai = d.__cuda_array_interface__
    assert cai["data"] == (ptr, False) and cai["shape"] == a.shape and np.dtype(cai["typestr"]) == a.dtype
    if a.dtype.kind != "u" or a.dtype.itemsize == 1:      # torch reads the interface for the signed types only
        assert torch.as_tensor(d, device="cuda").data_ptr() == ptr
    # import: a CAI-only producer and a DLPack producer, no copy, correct dtype, and the filter runs on it
    prod = _CaiProducer(d.tensor, a.dtype)
    for src in (prod, _DlpackProducer(d.tensor), d.tensor):
        v = b200.as_b200(src)
        assert isinstance(v, b200.B200Array) and v.dtype == a.dtype and v.shape == a.shape
        assert v.tensor.data_ptr() == ptr
        got = b200.ndfilters.uniform_filter(v, 3).to_numpy()
        assert_parity(got, ndi.uniform_filter(a, 3), exact=a.dtype.kind != "f")
    assert b200.B200Array.from_device_array(prod).tensor.data_ptr() == ptr
    with pytest.raises(TypeError):
        b200.B200Array.from_device_array(a)


@pytest.mark.parametrize("dt", ["float32", "float64", "uint16"])
def test_axes_subsets(b200, dt):
    """scipy's `axes=` (scipy/ndimage/_filters.py:516-553,1126-1139,1181-1195,1271-1292,1738-1830): derivative
    chains along the listed axes only, N-D windows expanded with singleton dimensions."""
    rng = np.random.default_rng(15)
    a = (rng.random((18, 21, 40)) * 200).astype(dt)
    exact = np.dtype(dt).kind != "f"
    d = dev(b200, a)
    for axes in [(0,), (2, 0), (1, 2), (2, 1, 0), ()]:
        k = len(axes)
        got = b200.ndimage.gaussian_gradient_magnitude(d, 1.2, axes=axes, mode="nearest")
        assert_parity(got.to_numpy(), ndi.gaussian_gradient_magnitude(a, 1.2, axes=axes, mode="nearest"), exact=exact,
                      data=a)
        got = b200.ndimage.gaussian_laplace(d, [0.8, 1.5, 1.1][:k], axes=axes)
        assert_parity(got.to_numpy(), ndi.gaussian_laplace(a, [0.8, 1.5, 1.1][:k], axes=axes), exact=exact, data=a)
        got = b200.ndimage.minimum_filter(d, size=[3, 2, 5][:k], axes=axes)
        assert_parity(got.to_numpy(), ndi.minimum_filter(a, size=[3, 2, 5][:k], axes=axes), exact=True)
        if k:
            w = rng.random((3, 4, 5)[:k]) - 0.3
            got = b200.ndimage.convolve(d, w, axes=axes, mode="wrap", origin=[1, -1, 0][:k])
            assert_parity(got.to_numpy(), ndi.convolve(a, w, axes=axes, mode="wrap", origin=[1, -1, 0][:k]),
                          exact=exact, data=a)
            fp = rng.random((3, 2, 3)[:k]) > 0.4
            fp.flat[0], fp.flat[-1] = True, False
            got = b200.ndimage.maximum_filter(d, footprint=fp, axes=axes, mode="mirror")
            assert_parity(got.to_numpy(), ndi.maximum_filter(a, footprint=fp, axes=axes, mode="mirror"), exact=True)


@pytest.mark.parametrize("name", ["laplace", "prewitt", "sobel"])
@pytest.mark.parametrize("dt", ["float64", "float32", "int16", "uint8"])
def test_three_tap_compositions(b200, name, dt):
    # SURVEY.md 8(f) rank 1; reference: tests/.../test__diff.py, test__edge.py
    a = (np.random.default_rng(6).random((24, 30)) * 100).astype(dt)
    d = b200.from_array(a, chunks=(12, 10))
    got = getattr(b200.ndfilters, name)(d)
    assert_parity(np.asarray(got), getattr(ndi, name)(a))
    if name != "laplace":
        got = getattr(b200.ndfilters, name)(d, axis=0, mode="mirror")
        assert_parity(np.asarray(got), getattr(ndi, name)(a, axis=0, mode="mirror"))


# ---------------------------------------------------------------------------
# threshold_local: the reference's stored known-answer
# masks (tests/test_dask_image/test_ndfilters/test__threshold.py:24-64)
# ---------------------------------------------------------------------------
_THRESH_IMAGE = np.array([[0, 0, 1, 3, 5], [0, 1, 4, 3, 4], [1, 2, 5, 4, 1], [2, 4, 5, 2, 1], [4, 5, 1, 0, 0]],
                         dtype=np.int64)
_THRESH_MASK = np.array([[False, False, False, False, True], [False, False, True, False, True],
                         [False, False, True, True, False], [False, True, True, False, False],
                         [True, True, False, False, False]])


@pytest.mark.parametrize("block_size", [3, [3, 3], np.array([3, 3])])
@pytest.mark.parametrize("chunked", [False, True])
def test_threshold_local_known_answers(b200, block_size, chunked):
    img = dev(b200, _THRESH_IMAGE)
    if chunked:
        img = b200.from_array(img, chunks=(3, 2))
    for kw in (dict(method="gaussian"), dict(method="gaussian", param=1. / 3.), dict(method="mean")):
        out = b200.ndfilters.threshold_local(img, block_size, **kw)
        assert np.asarray(out).dtype == np.float64
        assert np.array_equal(_THRESH_IMAGE > np.asarray(out), _THRESH_MASK), kw


def test_threshold_local_values_and_errors(b200):
    rng = np.random.default_rng(8)
    a = rng.integers(0, 255, (40, 48)).astype(np.uint8)
    got = b200.ndfilters.threshold_local(dev(b200, a), 7, offset=2.5, mode="mirror")
    ref = ndi.gaussian_filter(a.astype(np.float64), (7 - 1) / 6.0, mode="mirror") - 2.5
    assert_parity(np.asarray(got), ref)
    got = b200.ndfilters.threshold_local(dev(b200, a), (5, 3), method="mean", offset=-1.0)
    ref = ndi.uniform_filter(a.astype(np.float64), (5, 3)) + 1.0
    assert_parity(np.asarray(got), ref)
    with pytest.raises(ValueError):
        b200.ndfilters.threshold_local(dev(b200, a), 3, method="generic", param=None)
    with pytest.raises(ValueError):
        b200.ndfilters.threshold_local(dev(b200, a), 3, method="otsu")
    with pytest.raises(NotImplementedError):
        b200.ndfilters.threshold_local(dev(b200, a), 3, method="median")


# ---------------------------------------------------------------------------
# minimum / maximum filters: bit-identical to scipy for
# every dtype, separable boxes and general footprints, chunked or whole
# ---------------------------------------------------------------------------
EXT_DTYPES = [np.uint8, np.int16, np.uint16, np.int32, np.int64, np.uint64, np.float32, np.float64]


@pytest.mark.parametrize("dtype", EXT_DTYPES)
@pytest.mark.parametrize("mode", ["reflect", "mirror", "nearest", "constant", "wrap"])
def test_min_max_filters_bitwise(b200, dtype, mode):
    rng = np.random.default_rng(17)
    shape = (19, 14, 23)
    if np.dtype(dtype).kind == "f":
        a = (rng.random(shape) - 0.5).astype(dtype)
    elif np.dtype(dtype).itemsize == 8:
        a = rng.integers(0, 2 ** 62, shape).astype(dtype)       # beyond 2^53: the reference rounds through double
    else:
        info = np.iinfo(dtype)
        a = rng.integers(info.min, info.max, shape, endpoint=True).astype(dtype)
    d = dev(b200, a)
    fp = rng.random((3, 4, 5)) > 0.45
    fp[1, 2, 2] = True
    for name in ("minimum_filter", "maximum_filter"):
        f, sp = getattr(b200.ndimage, name), getattr(ndi, name)
        for kw in (dict(size=3), dict(size=(5, 1, 4), origin=(2, 0, -2)), dict(footprint=fp, origin=(0, 1, -1)),
                   dict(footprint=np.ones((2, 3, 2), bool))):
            got = f(d, mode=mode, cval=-7.5, **kw)
            assert_parity(got.to_numpy(), sp(a, mode=mode, cval=-7.5, **kw), exact=True, what="%s %s" % (name, kw))
        f1, sp1 = getattr(b200.ndimage, name + "1d"), getattr(ndi, name + "1d")
        for axis, size, origin in ((0, 4, -2), (1, 7, 3), (2, 2, 0), (-1, 30, 5)):
            got = f1(d, size, axis=axis, mode=mode, cval=300.7, origin=origin)
            assert_parity(got.to_numpy(), sp1(a, size, axis=axis, mode=mode, cval=300.7, origin=origin), exact=True)


def test_min_max_public_wrappers_and_errors(b200):
    rng = np.random.default_rng(3)
    a = rng.integers(0, 65536, (40, 36), dtype=np.uint16)
    d = dev(b200, a)
    fp = np.array([[0, 1, 0], [1, 1, 1], [0, 1, 0]], bool)
    for name in ("minimum_filter", "maximum_filter"):
        f, sp = getattr(b200.ndfilters, name), getattr(ndi, name)
        assert_parity(np.asarray(f(d, size=5)), sp(a, size=5), exact=True)
        assert_parity(np.asarray(f(d, footprint=fp, mode="wrap")), sp(a, footprint=fp, mode="wrap"), exact=True)
        chunked = f(b200.from_array(a, chunks=(16, 12)), size=(3, 7), origin=(1, -2))
        assert_parity(np.asarray(chunked), sp(a, size=(3, 7), origin=(1, -2)), exact=True)
        with pytest.raises(RuntimeError):
            f(d)                                    # neither size nor footprint
        with pytest.raises(RuntimeError):
            f(d, size=3, footprint=fp)              # both
        with pytest.raises(ValueError):
            f(d, size=3, origin=2)                  # origin outside the footprint
    with pytest.raises(ValueError):
        b200.ndimage.minimum_filter(d, footprint=np.zeros((3, 3), bool))
    assert b200.ndfilters.minimum_filter.__name__ == "minimum_filter"


def test_b200array_numpy_protocol_stays_on_device(b200):
    """NEP-18 surface dask needs to assemble halo'd blocks from B200Array chunks."""
    a = np.arange(24, dtype=np.float32).reshape(4, 6)
    d = dev(b200, a)
    c = np.concatenate([d[:, :2], d[:, 2:]], axis=1)
    assert type(c) is b200.B200Array and np.array_equal(np.asarray(c), a)
    s = np.stack([d, d], axis=0)
    assert type(s) is b200.B200Array and s.shape == (2, 4, 6)
    e = np.empty_like(d, shape=(3, 3))
    z = np.zeros_like(d)
    assert type(e) is b200.B200Array and e.shape == (3, 3) and e.dtype == np.float32
    assert type(z) is b200.B200Array and not np.asarray(z).any()
    assert np.shape(d) == (4, 6) and np.ndim(d) == 2
    assert np.array_equal(np.asarray(np.transpose(d)), a.T)
    with pytest.raises(TypeError):
        np.fft.fft(d)                 # not on the path: refused, not silently computed on the host
    # halo assembly the way map_overlap does it: neighbours' planes concatenated, filtered, trimmed
    top, bottom = d[:2], d[2:]
    block = np.concatenate([top, bottom[:1]], axis=0)
    got = b200.ndimage.uniform_filter(block, size=(3, 1), mode="nearest")[:2]
    ref = ndi.uniform_filter(a, size=(3, 1), mode="nearest")[:2]
    assert_parity(np.asarray(got)[:1], ref[:1])


def test_concurrent_calls_from_a_thread_pool(b200):
    """dask's threaded scheduler calls chunk functions concurrently (SURVEY.md 8b,
    'Threading'): the C ABI keeps only thread-local state, every call takes the
    calling thread's stream."""
    import threading
    from concurrent.futures import ThreadPoolExecutor

    import torch

    rng = np.random.default_rng(99)
    jobs = []
    for k in range(24):
        dt = [np.float32, np.uint16, np.float64][k % 3]
        a = (rng.random((20 + k, 24, 32)) * 1000).astype(dt)
        jobs.append((k, a))
    errors = []
    lock = threading.Lock()

    def work(job):
        k, a = job
        try:
            stream = torch.cuda.Stream()
            with torch.cuda.stream(stream):
                d = b200.B200Array.from_numpy(a)
                g = b200.ndfilters.gaussian_filter(d, sigma=1.0 + (k % 4) * 0.5)
                u = b200.ndfilters.uniform_filter(d, size=3 + (k % 3) * 2)
                m = b200.ndfilters.maximum_filter(d, size=3)
                stream.synchronize()
            assert_parity(np.asarray(g), ndi.gaussian_filter(a, 1.0 + (k % 4) * 0.5), data=a)
            assert_parity(np.asarray(u), ndi.uniform_filter(a, 3 + (k % 3) * 2), data=a)
            assert_parity(np.asarray(m), ndi.maximum_filter(a, size=3), exact=True)
        except Exception as exc:       # collected, not raised inside the pool
            with lock:
                errors.append("job %d: %r" % (k, exc))

    with ThreadPoolExecutor(8) as ex:
        list(ex.map(work, jobs))
    assert not errors, "\n".join(errors)
