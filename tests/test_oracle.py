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
