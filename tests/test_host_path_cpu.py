"""The host side of the single-array path, end to end on the CPU.

The product has no CPU fallback; here -- test infrastructure only -- the six raw
launch helpers of ``dask_image_b200.ndimage`` are replaced by the oracle (the C
ABI's contract restated on numpy views of CPU tensors) and ``B200Array`` is
allowed to wrap CPU tensors, so that everything ABOVE the C ABI runs for real:
argument handling, the cached plans and taps, pass chains, scratch ping-pong,
prefix sharing between chains, fused epilogues.  With exact arithmetic below
it, the public functions must reproduce scipy bit for bit, for every dtype.
"""
import numpy as np
import pytest
import scipy.ndimage as ndi
import torch

from _oracle_backend import _fold, _np
from dask_image_b200 import _array, _lib, ndimage
from oracle import oracle as orc


def _view(t, shape):
    return _np(t).reshape(tuple(shape))


@pytest.fixture
def cpu_path(monkeypatch):
    calls = []

    def init(self, tensor):
        self.tensor = tensor if tensor.is_contiguous() else tensor.contiguous()

    def corr1d(src, dst, shape, np_dtype, axis, weights, mode, cval, origin, pad=(0, 0),
               epilogue=_lib.EPI_STORE, flags=None, prepared=None):
        assert tuple(pad) == (0, 0)
        if prepared is not None:      # the cached view of the taps is the taps
            assert prepared[0].dtype == np.float64 and np.array_equal(prepared[0], np.asarray(weights, np.float64))
        calls.append(("corr", axis, len(weights), epilogue))
        r = orc.correlate1d(_view(src, shape), weights, axis, mode, cval, origin)
        _fold(_view(dst, shape), r.copy(), epilogue)

    def corr_pair(src, dst, shape, np_dtype, ps1, ps2, cval, epilogue=_lib.EPI_STORE, flags=None):
        # the fused launch IS the two passes, the intermediate stored in the array dtype
        if ps1.kind == "uniform":
            assert epilogue == _lib.EPI_STORE and np.dtype(np_dtype) in (np.uint8, np.uint16, np.float32)
            calls.append(("uniform_pair", ps1.axis, ps2.axis, ps1.size, ps2.size))
            mid = orc.uniform_filter1d(_view(src, shape), ps1.size, ps1.axis, ps1.mode, cval, ps1.origin)
            _view(dst, shape)[...] = orc.uniform_filter1d(mid, ps2.size, ps2.axis, ps2.mode, cval, ps2.origin)
            return
        calls.append(("pair", ps1.axis, ps2.axis, len(ps1.weights), len(ps2.weights), epilogue))
        mid = orc.correlate1d(_view(src, shape), ps1.weights, ps1.axis, ps1.mode, cval, ps1.origin)
        r = orc.correlate1d(mid, ps2.weights, ps2.axis, ps2.mode, cval, ps2.origin)
        _fold(_view(dst, shape), r.copy(), epilogue)

    def uniform1d(src, dst, shape, np_dtype, axis, size, mode, cval, origin, pad=(0, 0), flags=None):
        calls.append(("uniform", axis, size))
        _view(dst, shape)[...] = orc.uniform_filter1d(_view(src, shape), size, axis, mode, cval, origin)

    def corrnd(src, dst, shape, np_dtype, weights, origins, mode, cval, pad=(0, 0), flags=None):
        calls.append(("nd",))
        _view(dst, shape)[...] = orc.correlate(_view(src, shape), weights, mode, cval, list(origins))

    def minmax1d(src, dst, shape, np_dtype, axis, size, is_max, mode, cval, origin, pad=(0, 0), flags=None):
        calls.append(("minmax", axis, size))
        f = ndi.maximum_filter1d if is_max else ndi.minimum_filter1d
        _view(dst, shape)[...] = f(_view(src, shape), size, axis=axis, mode=mode, cval=cval, origin=origin)

    def minmaxnd(src, dst, shape, np_dtype, footprint, origins, is_max, mode, cval, pad=(0, 0), flags=None):
        calls.append(("minmax_nd",))
        f = ndi.maximum_filter if is_max else ndi.minimum_filter
        _view(dst, shape)[...] = f(_view(src, shape), footprint=footprint, mode=mode, cval=cval, origin=list(origins))

    def combine(acc, src, np_dtype, op):
        calls.append(("combine", op))
        _fold(_np(acc), _np(src).copy(), op)

    monkeypatch.setattr(_array.B200Array, "__init__", init)
    monkeypatch.setattr(_array, "default_device", lambda: torch.device("cpu"))
    for name, fn in [("_launch_correlate1d", corr1d), ("_launch_correlate1d_pair", corr_pair),
                     ("_launch_uniform1d", uniform1d),
                     ("_launch_correlate_nd", corrnd), ("_launch_minmax1d", minmax1d),
                     ("_launch_minmax_nd", minmaxnd), ("_launch_combine", combine)]:
        monkeypatch.setattr(ndimage, name, fn)
    ndimage._plan_cache.clear()
    return calls


DTYPES = [np.float32, np.float64, np.uint8, np.uint16, np.int16, np.int32]


def _data(rng, shape, dtype):
    dt = np.dtype(dtype)
    if dt.kind == "f":
        return (rng.random(shape) * 100 - 30).astype(dt)
    ii = np.iinfo(dt)
    return rng.integers(max(ii.min, -3000), min(ii.max, 3000), shape, endpoint=True).astype(dt)


def _same(got, ref):
    got = got.to_numpy()
    assert got.dtype == ref.dtype and got.shape == ref.shape
    np.testing.assert_array_equal(got, ref)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape", [(19, 23), (7, 9, 11)])
def test_public_functions_reproduce_scipy_bitwise(cpu_path, dtype, shape):
    rng = np.random.default_rng(17)
    a = _data(rng, shape, dtype)
    x = _array.B200Array.from_numpy(a)
    nd = len(shape)
    with np.errstate(all="ignore"):
        for _ in range(2):      # second round: every plan comes out of the cache
            _same(ndimage.gaussian_filter(x, 1.3), ndi.gaussian_filter(a, 1.3))
            _same(ndimage.gaussian_filter(x, (1.0, 0.0, 2.0)[:nd], order=(0, 1, 2)[:nd], mode="wrap"),
                  ndi.gaussian_filter(a, (1.0, 0.0, 2.0)[:nd], order=(0, 1, 2)[:nd], mode="wrap"))
            _same(ndimage.gaussian_filter(x, np.float32(0.9), mode="constant", cval=3.0, truncate=2.5),
                  ndi.gaussian_filter(a, np.float32(0.9), mode="constant", cval=3.0, truncate=2.5))
            _same(ndimage.uniform_filter(x, 3, mode="mirror"), ndi.uniform_filter(a, 3, mode="mirror"))
            _same(ndimage.uniform_filter(x, (4, 1, 5)[:nd], origin=(1, 0, -2)[:nd]),
                  ndi.uniform_filter(a, (4, 1, 5)[:nd], origin=(1, 0, -2)[:nd]))
            _same(ndimage.gaussian_gradient_magnitude(x, 1.1), ndi.gaussian_gradient_magnitude(a, 1.1))
            _same(ndimage.gaussian_gradient_magnitude(x, 0.8, mode="nearest", truncate=3.0),
                  ndi.gaussian_gradient_magnitude(a, 0.8, mode="nearest", truncate=3.0))
            _same(ndimage.gaussian_laplace(x, 1.2, mode="reflect"), ndi.gaussian_laplace(a, 1.2, mode="reflect"))
            _same(ndimage.laplace(x), ndi.laplace(a))
            _same(ndimage.sobel(x, axis=0), ndi.sobel(a, axis=0))
            _same(ndimage.prewitt(x, axis=-1, mode="constant", cval=1.0), ndi.prewitt(a, axis=-1, mode="constant", cval=1.0))
            w = rng.random((3,) * nd) - 0.3
            _same(ndimage.correlate(x, w, mode="wrap"), ndi.correlate(a, w, mode="wrap"))
            _same(ndimage.convolve(x, w, origin=(1, -1, 0)[:nd]), ndi.convolve(a, w, origin=(1, -1, 0)[:nd]))
            _same(ndimage.minimum_filter(x, size=3), ndi.minimum_filter(a, size=3))
            _same(ndimage.maximum_filter(x, footprint=np.eye(3, dtype=bool) if nd == 2 else None,
                                         size=None if nd == 2 else (1, 3, 2)),
                  ndi.maximum_filter(a, footprint=np.eye(3, dtype=bool) if nd == 2 else None,
                                     size=None if nd == 2 else (1, 3, 2)))
            w1 = rng.random(5) - 0.5
            _same(ndimage.correlate1d(x, w1, axis=0, origin=-1), ndi.correlate1d(a, w1, axis=0, origin=-1))
            _same(ndimage.gaussian_filter1d(x, 1.7, axis=-1, order=1), ndi.gaussian_filter1d(a, 1.7, axis=-1, order=1))


def test_gradient_magnitude_shares_chain_prefixes(cpu_path):
    """3-D gradient magnitude / laplace, unfused: 8 launches for scipy's 9 passes (the chains along the last
    two axes start with the same smoothing pass along axis 0), epilogues fused into each chain's last pass.
    With the fused pass pair (float32): 2 axis-0 passes + 3 pair launches."""
    a = np.random.default_rng(3).random((6, 7, 8)).astype(np.float32)
    x = _array.B200Array.from_numpy(a)
    old = ndimage.set_pair_fusion(False)
    try:
        del cpu_path[:]
        _same(ndimage.gaussian_gradient_magnitude(x, 1.0), ndi.gaussian_gradient_magnitude(a, 1.0))
        corr = [c for c in cpu_path if c[0] == "corr"]
        assert len(corr) == 8 and not [c for c in cpu_path if c[0] in ("combine", "pair")]
        assert [c[3] for c in corr if c[3] != _lib.EPI_STORE] == [_lib.EPI_SQ_STORE, _lib.EPI_SQ_ACC,
                                                                 _lib.EPI_SQ_ACC_SQRT]
        del cpu_path[:]
        _same(ndimage.gaussian_laplace(x, 1.0), ndi.gaussian_laplace(a, 1.0))
        corr = [c for c in cpu_path if c[0] == "corr"]
        assert len(corr) == 8
        assert [c[3] for c in corr if c[3] != _lib.EPI_STORE] == [_lib.EPI_ACC, _lib.EPI_ACC]
    finally:
        ndimage.set_pair_fusion(old)
    ndimage.set_pair_fusion(True)
    try:
        del cpu_path[:]
        _same(ndimage.gaussian_gradient_magnitude(x, 1.0), ndi.gaussian_gradient_magnitude(a, 1.0))
        assert [c[:3] for c in cpu_path] == [("corr", 0, 9), ("pair", 1, 2), ("corr", 0, 9), ("pair", 1, 2),
                                             ("pair", 1, 2)]
        assert [c[-1] for c in cpu_path if c[0] == "pair"] == [_lib.EPI_SQ_STORE, _lib.EPI_SQ_ACC,
                                                              _lib.EPI_SQ_ACC_SQRT]
        del cpu_path[:]
        _same(ndimage.gaussian_laplace(x, 1.0), ndi.gaussian_laplace(a, 1.0))
        assert [c[0] for c in cpu_path] == ["corr", "pair", "corr", "pair", "pair"]
        assert [c[-1] for c in cpu_path if c[0] == "pair"] == [_lib.EPI_STORE, _lib.EPI_ACC, _lib.EPI_ACC]
        # a 3-D gaussian is an axis-0 pass and one pair; a 2-D one is a single launch; float64 / integers and
        # filters without a fused kernel (anisotropic radii) stay on single passes
        del cpu_path[:]
        _same(ndimage.gaussian_filter(x, 1.0), ndi.gaussian_filter(a, 1.0))
        assert [c[0] for c in cpu_path] == ["corr", "pair"]
        del cpu_path[:]
        x2 = _array.B200Array.from_numpy(a[0])
        _same(ndimage.gaussian_filter(x2, 1.0), ndi.gaussian_filter(a[0], 1.0))
        assert [c[0] for c in cpu_path] == ["pair"]
        del cpu_path[:]
        _same(ndimage.gaussian_filter(x, (1.0, 1.0, 1.5)), ndi.gaussian_filter(a, (1.0, 1.0, 1.5)))
        assert [c[0] for c in cpu_path] == ["corr", "corr", "corr"]
        del cpu_path[:]
        x64 = _array.B200Array.from_numpy(a.astype(np.float64))
        _same(ndimage.gaussian_filter(x64, 1.0), ndi.gaussian_filter(a.astype(np.float64), 1.0))
        assert [c[0] for c in cpu_path] == ["corr", "corr", "corr"]
    finally:
        ndimage.set_pair_fusion(old)


def test_output_argument_and_input_untouched(cpu_path):
    a = np.random.default_rng(4).integers(0, 255, (12, 14)).astype(np.uint8)
    x = _array.B200Array.from_numpy(a)
    out = x.empty_like()
    r = ndimage.gaussian_filter(x, 2.0, output=out)
    assert r is out
    _same(out, ndi.gaussian_filter(a, 2.0))
    np.testing.assert_array_equal(x.to_numpy(), a)          # the input is never written
    # all sigmas ~ 0: identity copy, no launch
    del cpu_path[:]
    _same(ndimage.gaussian_filter(x, 0.0), a)
    assert cpu_path == []


@pytest.mark.parametrize("dtype", [np.float32, np.uint16])
def test_ndfilters_layer_chunked_equals_whole_array_scipy(cpu_path, dtype):
    """The mirror of dask_image.ndfilters over HaloChunked (the eager map_overlap stand-in): halo depths,
    boundary handling and trimming per ref:_gaussian.py:47-81, _smooth.py:17-40, _conv.py:15-66 --
    chunk-wise results are the whole-array scipy results, bit for bit."""
    import dask_image_b200 as b200

    rng = np.random.default_rng(23)
    a = _data(rng, (37, 41), dtype)
    with np.errstate(all="ignore"):
        for chunks in [(10, 13), (37, 41), (5, 40)]:
            d = b200.from_array(a, chunks=chunks)
            for mode in ("reflect", "constant", "wrap", "nearest", "mirror"):
                # (gaussian / uniform keep boundary="none" as the reference does, ref:_gaussian.py:70-81: 'wrap'
                # then wraps inside each edge chunk; only convolve / correlate rewrite it to a periodic halo)
                if mode != "wrap" or chunks == a.shape:
                    np.testing.assert_array_equal(
                        np.asarray(b200.ndfilters.gaussian_filter(d, 1.2, mode=mode).compute()),
                        ndi.gaussian_filter(a, 1.2, mode=mode), err_msg=str((chunks, mode)))
                    np.testing.assert_array_equal(
                        np.asarray(b200.ndfilters.uniform_filter(d, (3, 4), mode=mode, origin=(0, -1)).compute()),
                        ndi.uniform_filter(a, (3, 4), mode=mode, origin=(0, -1)), err_msg=str((chunks, mode)))
                w = rng.random((3, 4)) - 0.2
                np.testing.assert_array_equal(
                    np.asarray(b200.ndfilters.convolve(d, b200.B200Array.from_numpy(w), mode=mode).compute()),
                    ndi.convolve(a, w, mode=mode), err_msg=str((chunks, mode)))
            np.testing.assert_array_equal(np.asarray(b200.ndfilters.gaussian_gradient_magnitude(d, 1.0).compute()),
                                          ndi.gaussian_gradient_magnitude(a, 1.0))
            np.testing.assert_array_equal(np.asarray(b200.ndfilters.sobel(d, axis=1).compute()), ndi.sobel(a, axis=1))
            np.testing.assert_array_equal(np.asarray(b200.ndfilters.maximum_filter(d, size=3).compute()),
                                          ndi.maximum_filter(a, size=3))


@pytest.mark.parametrize("dtype", [np.uint8, np.uint16])
def test_integer_box_filter_pairs_the_last_two_axes(cpu_path, dtype):
    """uniform_filter on uint8 / uint16 with equal odd sizes, origin 0 and 16-byte rows: one axis-0 pass and ONE
    pair launch; anything else stays pass by pass.  Values are scipy's either way."""
    rng = np.random.default_rng(23)
    a = _data(rng, (6, 10, 32), dtype)
    x = _array.B200Array.from_numpy(a)
    del cpu_path[:]
    _same(ndimage.uniform_filter(x, 7), ndi.uniform_filter(a, 7))
    assert cpu_path == [("uniform", 0, 7), ("uniform_pair", 1, 2, 7, 7)]
    del cpu_path[:]
    _same(ndimage.uniform_filter(x, (1, 3, 3), mode="constant", cval=2.0),
          ndi.uniform_filter(a, (1, 3, 3), mode="constant", cval=2.0))
    assert cpu_path == [("uniform_pair", 1, 2, 3, 3)]
    for kw in (dict(size=4), dict(size=(3, 5, 7)), dict(size=5, origin=1), dict(size=19),
               dict(size=3, mode="constant", cval=0.5)):
        del cpu_path[:]
        _same(ndimage.uniform_filter(x, **kw), ndi.uniform_filter(a, **kw))
        assert all(c[0] == "uniform" for c in cpu_path) and len(cpu_path) == 3, (kw, cpu_path)
    # rows that are not 16-byte multiples
    b = np.ascontiguousarray(a[:, :, :27])
    del cpu_path[:]
    _same(ndimage.uniform_filter(_array.B200Array.from_numpy(b), 7), ndi.uniform_filter(b, 7))
    assert [c[0] for c in cpu_path] == ["uniform"] * 3
    old = ndimage.set_pair_fusion(False)
    try:
        del cpu_path[:]
        _same(ndimage.uniform_filter(x, 7), ndi.uniform_filter(a, 7))
        assert [c[0] for c in cpu_path] == ["uniform"] * 3
    finally:
        ndimage.set_pair_fusion(old)


def test_float_box_filter_pairs_the_last_two_axes(cpu_path):
    a = np.random.default_rng(29).random((6, 10, 19)).astype(np.float32)
    x = _array.B200Array.from_numpy(a)
    del cpu_path[:]
    _same(ndimage.uniform_filter(x, 5), ndi.uniform_filter(a, 5))
    assert cpu_path == [("uniform", 0, 5), ("uniform_pair", 1, 2, 5, 5)]
    del cpu_path[:]
    _same(ndimage.uniform_filter(x, (5, 4, 4)), ndi.uniform_filter(a, (5, 4, 4)))
    assert [c[0] for c in cpu_path] == ["uniform"] * 3
    d = _array.B200Array.from_numpy(a.astype(np.float64))
    del cpu_path[:]
    _same(ndimage.uniform_filter(d, 5), ndi.uniform_filter(a.astype(np.float64), 5))
    assert [c[0] for c in cpu_path] == ["uniform"] * 3


def test_chunk_functions_are_thread_safe(cpu_path):
    """dask's threaded scheduler calls the chunk function from a thread pool (SURVEY 8b "Threading"): the plan /
    taps / pass-grouping caches are shared, filled concurrently and wiped when full -- results must not depend on it."""
    from concurrent.futures import ThreadPoolExecutor

    rng = np.random.default_rng(41)
    chunks = [_data(rng, (9, 10, 11), dt) for dt in (np.float32, np.uint16, np.float64, np.int16) for _ in range(4)]
    sigmas = [0.6 + 0.1 * i for i in range(12)]

    def work(i):
        a = chunks[i % len(chunks)]
        x = _array.B200Array.from_numpy(a)
        s = sigmas[i % len(sigmas)]
        with np.errstate(all="ignore"):
            got = [ndimage.gaussian_filter(x, s).to_numpy(), ndimage.gaussian_laplace(x, s, mode="wrap").to_numpy(),
                   ndimage.uniform_filter(x, 3 + i % 3).to_numpy()]
            ref = [ndi.gaussian_filter(a, s), ndi.gaussian_laplace(a, s, mode="wrap"), ndi.uniform_filter(a, 3 + i % 3)]
        return all(np.array_equal(g, r) for g, r in zip(got, ref))

    old = ndimage._PLAN_CACHE_MAX, ndimage._TAPS_CACHE_MAX
    ndimage._PLAN_CACHE_MAX, ndimage._TAPS_CACHE_MAX = 8, 8      # force the caches to be wiped while in use
    try:
        ndimage._plan_cache.clear()
        ndimage._taps_cache.clear()
        with ThreadPoolExecutor(8) as pool:
            assert all(pool.map(work, range(96)))
    finally:
        ndimage._PLAN_CACHE_MAX, ndimage._TAPS_CACHE_MAX = old


def test_dlpack_export_is_zero_copy(cpu_path):
    """B200Array forwards the DLPack protocol of its buffer (consumers such as cupy / torch / jax view it in place);
    the CUDA Array Interface is only offered for device buffers."""
    x = _array.B200Array.from_numpy(np.arange(12, dtype=np.int16).reshape(3, 4))
    t = torch.from_dlpack(x)
    assert t.data_ptr() == x.tensor.data_ptr() and tuple(t.shape) == (3, 4) and t.dtype == torch.int16
    assert not hasattr(x, "__cuda_array_interface__")
    with pytest.raises(TypeError):
        _array.B200Array.from_device_array(np.zeros(3))
