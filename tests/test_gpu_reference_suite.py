"""The remaining value tests of the reference's ndfilters suite, run against the
CUDA path (``HaloChunked`` stands in for ``dask.array.from_array``; the expected
values come from scipy, as in the reference):

  test__gaussian.py:87-98 (shape/type), test__smooth.py:31-77 (shape/type,
  identity), test__conv.py:53-114 (shape/type, identity), test__diff.py:11-36,
  test__edge.py:42-92, test__order.py:68-210 (shape/type, identity, comprehensions,
  compare).
"""
import numpy as np
import pytest
import scipy.ndimage as ndi

from _helpers import assert_parity

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def b200():
    import dask_image_b200

    return dask_image_b200


def chunked(b200, a, chunks):
    return b200.from_array(b200.B200Array.from_numpy(np.asarray(a)), chunks=chunks)


A = np.arange(140.0).reshape(10, 14)


@pytest.mark.parametrize("call", [
    lambda f, d: f.gaussian(d, sigma=1.0, truncate=4.0),
    lambda f, d: f.gaussian_filter(d, sigma=1.0, truncate=4.0),
    lambda f, d: f.gaussian_gradient_magnitude(d, sigma=1.0, truncate=4.0),
    lambda f, d: f.gaussian_laplace(d, sigma=1.0, truncate=4.0),
    lambda f, d: f.uniform_filter(d, 1),
    lambda f, d: f.minimum_filter(d, size=1),
    lambda f, d: f.maximum_filter(d, footprint=np.ones((1, 1), bool)),
])
def test_shape_type(b200, call):
    d = chunked(b200, A, (5, 7))
    d2 = call(b200.ndfilters, d)
    assert all(type(s) is int for s in d2.shape) and d2.shape == A.shape and d2.dtype == A.dtype


def test_convolutions_shape_type_and_identity(b200):
    d = chunked(b200, A, (5, 7))
    w = b200.B200Array.from_numpy(np.ones((1, 1)))
    for name in ("convolve", "correlate"):
        d2 = getattr(b200.ndfilters, name)(d, w)
        assert all(type(s) is int for s in d2.shape)
        assert_parity(np.asarray(d2), A)
        assert_parity(np.asarray(d2), getattr(ndi, name)(A, np.ones((1, 1))))


def test_uniform_and_order_identity(b200):
    d = chunked(b200, A, (5, 7))
    assert_parity(np.asarray(b200.ndfilters.uniform_filter(d, 1, origin=0)), A)
    for name in ("minimum_filter", "maximum_filter"):
        for kw in (dict(size=1), dict(footprint=np.ones((1, 1)))):
            got = getattr(b200.ndfilters, name)(d, **kw)
            assert_parity(np.asarray(got), A, exact=True)
            assert_parity(np.asarray(got), getattr(ndi, name)(A, **kw), exact=True)


@pytest.mark.parametrize("name", ["minimum_filter", "maximum_filter"])
@pytest.mark.parametrize("size, footprint, origin", [
    (2, None, 0), (None, np.ones((2, 3)), 0), (None, np.ones((2, 3)), (0, 1)), (None, np.ones((2, 3)), (0, -1)),
    (None, (np.mgrid[-2:3, -2:3] ** 2).sum(axis=0) < 2.5 ** 2, 0),
    (None, (np.mgrid[-2:3, -2:3] ** 2).sum(axis=0) < 2.5 ** 2, (1, 2)),
    (None, (np.mgrid[-2:3, -2:3] ** 2).sum(axis=0) < 2.5 ** 2, (-1, -2)),
    (5, None, 0), (7, None, 0), (8, None, 0), (10, None, 0), (5, None, 2), (5, None, -2),
])
def test_ordered_filter_compare(b200, name, size, footprint, origin):
    # reference: test__order.py:150-210
    a = np.arange(140.0).reshape(10, 14)
    for chunks in ((5, 7), (10, 14)):
        got = getattr(b200.ndfilters, name)(chunked(b200, a, chunks), size=size, footprint=footprint, origin=origin)
        assert_parity(np.asarray(got), getattr(ndi, name)(a, size=size, footprint=footprint, origin=origin),
                      exact=True)


@pytest.mark.parametrize("name, kw", [("minimum_filter", dict(size=1)), ("maximum_filter", dict(size=3)),
                                      ("laplace", {}), ("prewitt", {}), ("sobel", {})])
def test_comprehensions(b200, name, kw):
    """Filtering the members of a stack one by one (reference: *_comprehensions)."""
    rng = np.random.default_rng(0)
    a = rng.random((3, 12, 14))
    d = chunked(b200, a, (3, 6, 7))
    f, sp = getattr(b200.ndfilters, name), getattr(ndi, name)
    exact = name in ("minimum_filter", "maximum_filter")
    for i in range(len(a)):
        assert_parity(np.asarray(f(d[i], **kw)), sp(a[i], **kw), exact=exact)


def test_laplace_compare(b200):
    a = np.arange(140.0).reshape(10, 14)
    assert_parity(np.asarray(b200.ndfilters.laplace(chunked(b200, a, (5, 7)))), ndi.laplace(a))


@pytest.mark.parametrize("name", ["prewitt", "sobel"])
@pytest.mark.parametrize("axis", [0, 1, 2, -1, -2, -3])
def test_edge_func_compare(b200, name, axis):
    # reference: test__edge.py:58-92
    a = np.arange(140.0 * 3).reshape(3, 10, 14) % 17
    got = getattr(b200.ndfilters, name)(chunked(b200, a, (3, 5, 7)), axis)
    assert_parity(np.asarray(got), getattr(ndi, name)(a, axis))
