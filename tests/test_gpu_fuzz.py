"""Randomised GPU parity (hypothesis, derandomised so every run sees the same
cases): the public per-chunk functions against the oracle over random shapes
(TMA-eligible and not), dtypes, modes, origins, tap counts and sigmas."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from _helpers import assert_parity
from oracle import oracle as orc

pytestmark = pytest.mark.gpu

import os

_N = int(os.environ.get("B200_FUZZ_EXAMPLES", "0"))      # raise for a longer hunt
_DT = ["uint8", "uint16", "int16", "int32", "float32", "float64"]
_MODES = ["nearest", "wrap", "reflect", "mirror", "constant"]


@pytest.fixture(scope="module")
def b200():
    import dask_image_b200

    return dask_image_b200


def _array(data, dt, max_nd=3):
    nd = data.draw(st.integers(1, max_nd))
    # last extent often a multiple of 16 so that the tiled kernels are taken
    last = data.draw(st.sampled_from([16, 32, 48, 64, 80, 7, 21, 33]))
    shape = tuple(data.draw(st.integers(1, 40)) for _ in range(nd - 1)) + (last,)
    rng = np.random.default_rng(data.draw(st.integers(0, 2 ** 16)))
    a = (rng.random(shape) * 200 - (60 if dt.kind != "u" else 0)).astype(dt)
    return a, rng


@settings(max_examples=_N or 80, deadline=None, derandomize=True)
@given(st.data())
def test_fuzz_1d_primitives(b200, data):
    dt = np.dtype(data.draw(st.sampled_from(_DT)))
    a, rng = _array(data, dt)
    mode = data.draw(st.sampled_from(_MODES))
    cval = data.draw(st.sampled_from([0.0, 7.0, -3.5]))
    axis = data.draw(st.integers(0, a.ndim - 1))
    nw = data.draw(st.integers(1, 40))
    origin = data.draw(st.integers(-(nw // 2), (nw - 1) // 2))
    w = rng.random(nw) - 0.4
    d = b200.B200Array.from_numpy(a)
    got = b200.ndimage.correlate1d(d, w, axis=axis, mode=mode, cval=cval, origin=origin)
    assert_parity(got.to_numpy(), orc.correlate1d(a, w, axis, mode, cval, origin), what="correlate1d", data=a)
    got = b200.ndimage.uniform_filter1d(d, nw, axis=axis, mode=mode, cval=cval, origin=origin)
    assert_parity(got.to_numpy(), orc.uniform_filter1d(a, nw, axis, mode, cval, origin), what="uniform1d", data=a)
    import scipy.ndimage as ndi

    for name in ("minimum_filter1d", "maximum_filter1d"):
        got = getattr(b200.ndimage, name)(d, nw, axis=axis, mode=mode, cval=cval, origin=origin)
        ref = getattr(ndi, name)(a, nw, axis=axis, mode=mode, cval=cval, origin=origin)
        assert_parity(got.to_numpy(), ref, exact=True, what=name)


@settings(max_examples=_N or 50, deadline=None, derandomize=True)
@given(st.data())
def test_fuzz_drivers(b200, data):
    dt = np.dtype(data.draw(st.sampled_from(_DT)))
    a, rng = _array(data, dt)
    mode = data.draw(st.sampled_from(_MODES))
    nd = a.ndim
    sigma = tuple(data.draw(st.sampled_from([0.0, 0.7, 1.0, 2.0, 3.3])) for _ in range(nd))
    order = tuple(data.draw(st.integers(0, 2)) for _ in range(nd))
    d = b200.B200Array.from_numpy(a)
    with np.errstate(all="ignore"):
        got = b200.ndimage.gaussian_filter(d, sigma, order=order, mode=mode, cval=1.5)
        assert_parity(got.to_numpy(), orc.gaussian_filter(a, sigma, order, mode, 1.5), what="gaussian", data=a)
        got = b200.ndimage.gaussian_gradient_magnitude(d, sigma, mode=mode, cval=1.5)
        assert_parity(got.to_numpy(), orc.gaussian_gradient_magnitude(a, sigma, mode, 1.5), what="ggm", data=a)
        got = b200.ndimage.gaussian_laplace(d, sigma, mode=mode, cval=1.5)
        assert_parity(got.to_numpy(), orc.gaussian_laplace(a, sigma, mode, 1.5), what="laplace", data=a)
    size = tuple(data.draw(st.integers(1, 9)) for _ in range(nd))
    origin = tuple(data.draw(st.integers(-(s // 2), (s - 1) // 2)) for s in size)
    got = b200.ndimage.uniform_filter(d, size, mode=mode, cval=1.5, origin=origin)
    assert_parity(got.to_numpy(), orc.uniform_filter(a, size, mode, 1.5, origin), what="uniform", data=a)
    if nd >= 2:
        wshape = tuple(data.draw(st.integers(1, 5)) for _ in range(nd))
        wn = rng.random(wshape) - 0.4
        origins = tuple(data.draw(st.integers(-(s // 2), (s - 1) // 2)) for s in wshape)
        got = b200.ndimage.correlate(d, wn, mode=mode, cval=1.5, origin=origins)
        assert_parity(got.to_numpy(), orc.correlate(a, wn, mode, 1.5, origins), what="correlate", data=a)
        got = b200.ndimage.convolve(d, wn, mode=mode, cval=1.5, origin=origins)
        assert_parity(got.to_numpy(), orc.convolve(a, wn, mode, 1.5, origins), what="convolve", data=a)
