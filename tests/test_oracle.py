"""CPU tests: the oracle (oracle/) pinned against scipy.ndimage -- the
arithmetic the reference runs per chunk -- and against the committed golden
vectors generated from it (tests/golden/make_golden.py)."""
import numpy as np
import pytest
import scipy.ndimage as ndi

from _helpers import golden_case, golden_cases
from oracle import oracle as orc

CASES = golden_cases()


@pytest.mark.parametrize("i", range(len(CASES)), ids=[c["name"] for c in CASES])
def test_oracle_matches_golden(i):
    func, a, kwargs, expected = golden_case(i)
    got = getattr(orc, func)(a, **kwargs)
    assert got.dtype == expected.dtype
    # the restatement is bit-exact for every dtype, floats included
    assert np.array_equal(got, expected, equal_nan=True)


def test_scipy_docstring_known_answers():
    # scipy/ndimage/_filters.py:582-583, 1528-1529, 728-731
    assert orc.correlate1d(np.array([2, 8, 0, 4, 1, 9, 9, 0]), [1, 3]).tolist() == \
        [8, 26, 8, 12, 7, 28, 36, 9]
    assert orc.uniform_filter1d(np.array([2, 8, 0, 4, 1, 9, 9, 0]), size=3).tolist() == \
        [4, 3, 4, 1, 4, 6, 6, 3]
    np.testing.assert_allclose(
        orc.gaussian_filter1d(np.array([1.0, 2.0, 3.0, 4.0, 5.0]), 1),
        [1.42704095, 2.06782203, 3.0, 3.93217797, 4.57295905], atol=1e-8)
    np.testing.assert_allclose(
        orc.gaussian_filter1d(np.array([1.0, 2.0, 3.0, 4.0, 5.0]), 4),
        [2.91948343, 2.95023502, 3.0, 3.04976498, 3.08051657], atol=1e-8)


@pytest.mark.parametrize("dt", ["uint8", "int16", "uint16", "int32", "uint64", "float32", "float64"])
@pytest.mark.parametrize("mode", ["nearest", "wrap", "reflect", "mirror", "constant"])
def test_oracle_matches_live_scipy(dt, mode):
    rng = np.random.default_rng(7)
    a = (rng.random((6, 1, 9)) * 90).astype(dt)
    for ax in range(3):
        for w, origin in [(rng.random(5) - 0.5, 0), (rng.random(4), -2), (rng.random(13), 3)]:
            assert np.array_equal(ndi.correlate1d(a, w, axis=ax, mode=mode, cval=2.5, origin=origin),
                                  orc.correlate1d(a, w, ax, mode, 2.5, origin))
        for size, origin in [(3, 0), (6, -3), (15, 2)]:
            assert np.array_equal(ndi.uniform_filter1d(a, size, axis=ax, mode=mode, cval=2.5, origin=origin),
                                  orc.uniform_filter1d(a, size, ax, mode, 2.5, origin))
    w = rng.random((3, 1, 4)) - 0.3
    assert np.array_equal(ndi.correlate(a, w, mode=mode, cval=2.5, origin=(1, 0, -2)),
                          orc.correlate(a, w, mode, 2.5, (1, 0, -2)))
    assert np.array_equal(ndi.convolve(a, w, mode=mode, cval=2.5, origin=(-1, 0, 1)),
                          orc.convolve(a, w, mode, 2.5, (-1, 0, 1)))


def test_oracle_errors():
    a = np.zeros((4, 4))
    with pytest.raises(ValueError):
        orc.correlate1d(a, [1, 1, 1], origin=2)
    with pytest.raises(RuntimeError):
        orc.correlate1d(a, [1, 1, 1], mode="bogus")
    with pytest.raises(RuntimeError):
        orc.correlate1d(a, [])
    with pytest.raises(RuntimeError):
        orc.uniform_filter1d(a, 0)


@pytest.mark.parametrize("mode", ["reflect", "constant", "mirror", "nearest"])
def test_map_overlap_emulation_equals_whole_array(mode):
    """What the reference computes chunk-wise (halo = ceil(sigma*truncate),
    boundary='none', dask_image/ndfilters/_gaussian.py:47-81) equals scipy on
    the whole array -- the reason whole-array scipy is the oracle."""
    a = np.random.default_rng(3).random((60, 66))
    whole = ndi.gaussian_filter(a, 2.0, mode=mode)
    chunked = orc.map_overlap_emulated(ndi.gaussian_filter, a, (30, 33), orc.gaussian_depth(2.0, 4.0, 2),
                                       "none", n_threads=2, sigma=2.0, mode=mode)
    assert np.array_equal(whole, chunked)


def test_map_overlap_emulation_periodic_wrap():
    # dask_image/ndfilters/_conv.py:23-25
    rng = np.random.default_rng(4)
    a, w = rng.random((40, 44)), rng.random((5, 4))
    whole = ndi.convolve(a, w, mode="wrap")
    chunked = orc.map_overlap_emulated(ndi.convolve, a, (20, 22), (2, 2), "periodic", weights=w, mode="constant")
    assert np.array_equal(whole, chunked)


# ---------------------------------------------------------------------------
# randomised pinning against live scipy (bounded hypothesis run)
# ---------------------------------------------------------------------------
from hypothesis import given, settings, strategies as st  # noqa: E402

_DT = ["uint8", "int8", "uint16", "int16", "uint32", "int32", "int64", "float32", "float64"]
_MODES = ["nearest", "wrap", "reflect", "mirror", "constant"]


@settings(max_examples=120, deadline=None)
@given(st.data())
def test_oracle_fuzz_against_scipy(data):
    nd = data.draw(st.integers(1, 3))
    shape = tuple(data.draw(st.integers(1, 9)) for _ in range(nd))
    dt = np.dtype(data.draw(st.sampled_from(_DT)))
    mode = data.draw(st.sampled_from(_MODES))
    cval = data.draw(st.sampled_from([0.0, -3.5, 2.25, 300.0]))
    seed = data.draw(st.integers(0, 2 ** 16))
    rng = np.random.default_rng(seed)
    a = (rng.random(shape) * 200 - (60 if dt.kind != "u" else 0)).astype(dt)
    axis = data.draw(st.integers(0, nd - 1))
    nw = data.draw(st.integers(1, 12))
    origin = data.draw(st.integers(-(nw // 2), (nw - 1) // 2))
    w = rng.random(nw) - 0.4
    if data.draw(st.booleans()) and nw % 2:
        w = (w + w[::-1]) / 2 if data.draw(st.booleans()) else (w - w[::-1]) / 2     # the folded branches
    assert np.array_equal(ndi.correlate1d(a, w, axis=axis, mode=mode, cval=cval, origin=origin),
                          orc.correlate1d(a, w, axis, mode, cval, origin), equal_nan=True)
    assert np.array_equal(ndi.uniform_filter1d(a, nw, axis=axis, mode=mode, cval=cval, origin=origin),
                          orc.uniform_filter1d(a, nw, axis, mode, cval, origin), equal_nan=True)
    wshape = tuple(data.draw(st.integers(1, 4)) for _ in range(nd))
    wn = rng.random(wshape) - 0.4
    origins = tuple(data.draw(st.integers(-(s // 2), (s - 1) // 2)) for s in wshape)
    assert np.array_equal(ndi.correlate(a, wn, mode=mode, cval=cval, origin=origins),
                          orc.correlate(a, wn, mode, cval, origins), equal_nan=True)
    conv_origins = tuple(data.draw(st.integers(-(s // 2), (s - 1) // 2)) for s in wshape)
    assert np.array_equal(ndi.convolve(a, wn, mode=mode, cval=cval, origin=conv_origins),
                          orc.convolve(a, wn, mode, cval, conv_origins), equal_nan=True)


@settings(max_examples=60, deadline=None)
@given(st.data())
def test_oracle_driver_fuzz_against_scipy(data):
    nd = data.draw(st.integers(1, 3))
    shape = tuple(data.draw(st.integers(1, 8)) for _ in range(nd))
    dt = np.dtype(data.draw(st.sampled_from(_DT)))
    mode = data.draw(st.sampled_from(_MODES))
    rng = np.random.default_rng(data.draw(st.integers(0, 2 ** 16)))
    a = (rng.random(shape) * 100).astype(dt)
    sigma = tuple(data.draw(st.sampled_from([0.0, 0.6, 1.0, 1.7])) for _ in range(nd))
    order = tuple(data.draw(st.integers(0, 2)) for _ in range(nd))
    truncate = data.draw(st.sampled_from([1.0, 2.5, 4.0]))
    with np.errstate(all="ignore"):
        assert np.array_equal(ndi.gaussian_filter(a, sigma, order=order, mode=mode, cval=1.5, truncate=truncate),
                              orc.gaussian_filter(a, sigma, order, mode, 1.5, truncate), equal_nan=True)
        assert np.array_equal(ndi.gaussian_gradient_magnitude(a, sigma, mode=mode, cval=1.5, truncate=truncate),
                              orc.gaussian_gradient_magnitude(a, sigma, mode, 1.5, truncate), equal_nan=True)
        assert np.array_equal(ndi.gaussian_laplace(a, sigma, mode=mode, cval=1.5, truncate=truncate),
                              orc.gaussian_laplace(a, sigma, mode, 1.5, truncate), equal_nan=True)
    size = tuple(data.draw(st.integers(1, 5)) for _ in range(nd))
    origin = tuple(data.draw(st.integers(-(s // 2), (s - 1) // 2)) for s in size)
    assert np.array_equal(ndi.uniform_filter(a, size, mode=mode, cval=1.5, origin=origin),
                          orc.uniform_filter(a, size, mode, 1.5, origin), equal_nan=True)
