"""CPU tests of the host side: argument validation tables of the reference's
tests (tests/test_dask_image/test_ndfilters/test__utils.py:63-189,
test__gaussian.py:12-35, test__conv.py:19-43, test__smooth.py:12-28), the
dispatch mechanism, the gaussian tap generator and the C-ABI library's export
table.  No kernel is launched here."""
import ctypes
import inspect
import os
import re

import numpy as np
import pytest
import scipy.ndimage

import dask_image_b200
from dask_image_b200 import _lib, ndimage
from dask_image_b200.dispatch import Dispatcher, check_arraytypes_compatible, get_type
from dask_image_b200.ndfilters import _gaussian, _utils

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---- C ABI -----------------------------------------------------------------
def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "b200_ndfilters.h")).read()
    declared = set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", header))
    assert {"b200_correlate1d", "b200_uniform_filter1d", "b200_correlate_nd", "b200_combine"} <= declared
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), "missing export %s" % name
    assert declared == set(_lib.EXPORTS)
    assert "sm_100a" in _lib.version()


def test_abi_argument_errors_without_gpu():
    """Validation happens before any CUDA call, with scipy's exception classes."""
    lib = _lib.load()
    shape = _lib.i64_array([4, 4])
    w = np.ones(3)
    fake = ctypes.c_void_p(256)
    fake2 = ctypes.c_void_p(512)
    rc = lib.b200_correlate1d(fake, fake2, 9, 2, shape, 0, _lib.f64_pointer(w), 3, 2, 0.0, 2, 0, 0, 0, 0, None)
    assert rc == _lib.EORIGIN
    with pytest.raises(ValueError, match="Invalid origin"):
        _lib.check(rc)
    rc = lib.b200_correlate1d(fake, fake2, 9, 2, shape, 0, _lib.f64_pointer(w), 3, 7, 0.0, 0, 0, 0, 0, 0, None)
    with pytest.raises(RuntimeError, match="boundary mode not supported"):
        _lib.check(rc)
    rc = lib.b200_correlate1d(fake, fake2, 42, 2, shape, 0, _lib.f64_pointer(w), 3, 2, 0.0, 0, 0, 0, 0, 0, None)
    with pytest.raises(RuntimeError, match="data type not supported"):
        _lib.check(rc)
    rc = lib.b200_correlate1d(fake, fake2, 9, 2, shape, 0, _lib.f64_pointer(w), 0, 2, 0.0, 0, 0, 0, 0, 0, None)
    with pytest.raises(RuntimeError, match="no filter weights given"):
        _lib.check(rc)
    rc = lib.b200_uniform_filter1d(fake, fake2, 9, 2, shape, 0, 0, 2, 0.0, 0, 0, 0, 0, None)
    with pytest.raises(RuntimeError, match="incorrect filter size"):
        _lib.check(rc)
    rc = lib.b200_uniform_filter1d(fake, fake2, 9, 2, shape, 0, 3, 2, 0.0, 2, 0, 0, 0, None)
    with pytest.raises(ValueError, match="invalid origin"):
        _lib.check(rc)
    rc = lib.b200_correlate1d(fake, fake, 9, 2, shape, 0, _lib.f64_pointer(w), 3, 2, 0.0, 0, 0, 0, 0, 0, None)
    with pytest.raises(RuntimeError, match="alias"):
        _lib.check(rc)
    # empty arrays are a no-op
    rc = lib.b200_correlate1d(None, None, 9, 2, _lib.i64_array([0, 4]), 0, _lib.f64_pointer(w), 3, 2, 0.0, 0,
                              0, 0, 0, 0, None)
    assert rc == 0
    with pytest.raises(RuntimeError, match="boundary mode not supported"):
        _lib.mode_code("bogus")


def test_plain_c_program_uses_the_abi(tmp_path):
    """The boundary is a C ABI: the header is C99 (no C++ / torch types in any signature), a C program links
    against the library and gets the documented status codes from argument validation -- no Python, no GPU
    (tests/native/abi_check.c)."""
    import shutil
    import subprocess

    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    _lib.load()
    src = os.path.join(ROOT, "tests", "native", "abi_check.c")
    exe = str(tmp_path / "abi_check")
    subprocess.check_call([gcc, "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic",
                           "-I", os.path.join(ROOT, "include"), src, "-o", exe, _lib.LIB_PATH,
                           "-Wl,-rpath," + os.path.dirname(_lib.LIB_PATH)])
    run = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert run.returncode == 0 and "abi_check: ok" in run.stdout, run.stdout + run.stderr


# ---- gaussian taps -----------------------------------------------------------
@pytest.mark.parametrize("order", [0, 1, 2, 3, 4])
@pytest.mark.parametrize("sigma", [0.5, 1.0, 2.0, 3.0, 4.0, 7.3])
def test_gaussian_taps_bit_identical_to_scipy(sigma, order):
    for truncate in (0.0, 2.0, 4.0):
        radius = int(truncate * sigma + 0.5)
        ref = scipy.ndimage._filters._gaussian_kernel1d(sigma, order, radius)
        got = ndimage._gaussian_kernel1d(sigma, order, radius)
        assert np.array_equal(ref, got)
        assert np.array_equal(ref[::-1], ndimage._gaussian_weights(sigma, order, truncate))


def test_gaussian_taps_cache_keeps_sigma_type_and_is_read_only():
    """scipy squares sigma in sigma's own type: a float32 scalar and the equal Python float give
    different taps, so they are different cache entries; cached taps cannot be modified."""
    ndimage._taps_cache.clear()
    s32, s64 = np.float32(2.1), float(np.float32(2.1))
    for _ in range(2):          # second round is served from the cache
        w32 = ndimage._gaussian_weights(s32, 0, 4.0)
        w64 = ndimage._gaussian_weights(s64, 0, 4.0)
        assert np.array_equal(w32, scipy.ndimage._filters._gaussian_kernel1d(s32, 0, 8)[::-1])
        assert np.array_equal(w64, scipy.ndimage._filters._gaussian_kernel1d(s64, 0, 8)[::-1])
        assert not np.array_equal(w32, w64)
    assert ndimage._gaussian_weights(s64, 0, 4.0) is w64
    assert ndimage._gaussian_weights(s64, 1, 4.0) is not w64
    assert ndimage._gaussian_weights(s64, 0, 4.0, radius=3).shape == (7,)
    with pytest.raises(ValueError):
        w64[0] = 1.0
    with pytest.raises(ValueError):
        ndimage._gaussian_weights(2.0, 0, 4.0, radius=-1)
    with pytest.raises(ValueError):
        ndimage._gaussian_weights(2.0, -1, 4.0)


def test_plan_cache_is_keyed_by_value_and_type():
    ndimage._plan_cache.clear()
    p1 = ndimage._cached_plan(ndimage._plan_gaussian_filter, 3, (2.0, 0, "reflect", 0.0, 4.0, None, None), {})
    p2 = ndimage._cached_plan(ndimage._plan_gaussian_filter, 3, (2.0, 0, "reflect", 0.0, 4.0, None, None), {})
    assert p1 is p2 and len(p1.chains[0][0]) == 3
    # other values, other types (np.float32 sigma, a tuple, an int), other rank: other plans
    others = [
        ndimage._cached_plan(ndimage._plan_gaussian_filter, 3, (2.5, 0, "reflect", 0.0, 4.0, None, None), {}),
        ndimage._cached_plan(ndimage._plan_gaussian_filter, 3, (np.float32(2.0), 0, "reflect", 0.0, 4.0, None, None), {}),
        ndimage._cached_plan(ndimage._plan_gaussian_filter, 3, ((2.0, 2.0, 2.0), 0, "reflect", 0.0, 4.0, None, None), {}),
        ndimage._cached_plan(ndimage._plan_gaussian_filter, 3, (2, 0, "reflect", 0.0, 4.0, None, None), {}),
        ndimage._cached_plan(ndimage._plan_gaussian_filter, 2, (2.0, 0, "reflect", 0.0, 4.0, None, None), {}),
        ndimage._cached_plan(ndimage._plan_gaussian_filter, 3, (2.0, 0, "mirror", 0.0, 4.0, None, None), {}),
    ]
    assert all(o is not p1 for o in others) and len({id(o) for o in others}) == len(others)
    # the cached plan is what a fresh build gives
    fresh = ndimage._plan_gaussian_filter(3, 2.0, 0, "reflect", 0.0, 4.0, None, None)
    assert [ps.sig for ps in fresh.chains[0][0]] == [ps.sig for ps in p1.chains[0][0]]
    # keyword arguments take part; unhashable arguments (arrays) bypass the cache
    k1 = ndimage._cached_plan(ndimage._plan_gaussian_gradient_magnitude, 2, (1.0, "reflect", 0.0, None), {"truncate": 3.0})
    k2 = ndimage._cached_plan(ndimage._plan_gaussian_gradient_magnitude, 2, (1.0, "reflect", 0.0, None), {"truncate": 2.0})
    assert k1 is not k2
    n = len(ndimage._plan_cache)
    a1 = ndimage._cached_plan(ndimage._plan_uniform_filter, 2, (np.array([3, 5]),), {})
    a2 = ndimage._cached_plan(ndimage._plan_uniform_filter, 2, (np.array([3, 5]),), {})
    assert a1 is not a2 and len(ndimage._plan_cache) == n
    # invalid parameters raise every time and are not cached
    for _ in range(2):
        with pytest.raises(RuntimeError):
            ndimage._cached_plan(ndimage._plan_gaussian_filter, 2, (1.0, 0, "bogus", 0.0, 4.0, None, None), {})


def test_pass_prepares_signature_and_taps_once():
    w = np.array([1, 2, 1], dtype=np.float32)          # any real dtype in, float64 taps out
    ps = ndimage._Pass("corr", 1, "reflect", weights=w, origin=0)
    assert ps.prepared[0].dtype == np.float64 and np.array_equal(ps.prepared[0], [1.0, 2.0, 1.0])
    assert ndimage._pass_signature(ps) == ("corr", 1, "reflect", 0, 0, np.array([1.0, 2.0, 1.0]).tobytes())
    box = ndimage._Pass("uniform", 0, "wrap", size=5, origin=1)
    assert box.prepared is None and ndimage._pass_signature(box) == ("uniform", 0, "wrap", 1, 5, None)


# ---- wrapper metadata (reference test__utils.py:11-60) -----------------------
def test__get_docstring_and_update_wrapper():
    f = lambda: 0  # noqa: E731
    expected = """
    Wrapped copy of "{mod_name}.{func_name}"


    Excludes the output parameter as it would not work with Dask arrays.


    Original docstring:

    {doc}
    """.format(mod_name=inspect.getmodule(f).__name__, func_name=f.__name__, doc="")
    assert _utils._get_docstring(f) == expected

    @_utils._update_wrapper(f)
    def g():
        return f()

    assert g.__name__ == f.__name__
    assert g.__doc__ == expected


def test_public_surface():
    nf = dask_image_b200.ndfilters
    for name in ["convolve", "correlate", "gaussian", "gaussian_filter", "gaussian_gradient_magnitude",
                 "gaussian_laplace", "uniform_filter", "laplace", "prewitt", "sobel"]:
        f = getattr(nf, name)
        assert f.__module__ == "dask_image_b200.ndfilters"
    assert 'Wrapped copy of "scipy.ndimage' in nf.gaussian_filter.__doc__

    def fake(input, output=None):
        """Summary of input.

        Parameters
        ----------
    input : array_like
        The input array.
    output : array, optional
        Dropped.
    mode : str
        Kept.

        Examples
        --------
        >>> gone
        """

    doc = _utils._get_docstring(fake)
    assert "output :" not in doc and "Dropped" not in doc and "gone" not in doc
    assert "image : array_like" in doc and "mode : str" in doc
    sig = inspect.signature(nf.gaussian_filter)
    assert list(sig.parameters) == ["image", "sigma", "order", "mode", "cval", "truncate"]
    assert list(inspect.signature(nf.uniform_filter).parameters) == ["image", "size", "mode", "cval", "origin"]
    assert list(inspect.signature(nf.convolve).parameters) == ["image", "weights", "mode", "cval", "origin"]


# ---- validation tables -------------------------------------------------------
@pytest.mark.parametrize("err_type, ndim, depth, boundary", [
    (TypeError, lambda: 0, 1, None), (TypeError, 1.0, 1, None), (ValueError, -1, 1, None),
    (TypeError, 1, lambda: 0, None), (TypeError, 1, 1.0, None), (ValueError, 1, -1, None),
    (ValueError, 1, (1, 1), None), (ValueError, 1, {0: 1, 1: 1}, None), (TypeError, 1, {1}, None),
    (TypeError, 1, 1, 1), (ValueError, 1, 1, (None, None)), (ValueError, 1, 1, {0: None, 1: None}),
    (TypeError, 1, 1, (1,)), (TypeError, 1, 1, {1}),
])
def test_errs__get_depth_boundary(err_type, ndim, depth, boundary):
    with pytest.raises(err_type):
        _utils._get_depth_boundary(ndim, depth, boundary)


@pytest.mark.parametrize("err_type, ndim, size", [
    (TypeError, 1.0, 1), (RuntimeError, 1, [[1]]), (TypeError, 1, 1.0), (TypeError, 1, [1.0]),
    (RuntimeError, 1, [1, 1]),
])
def test_errs__get_size(err_type, ndim, size):
    with pytest.raises(err_type):
        _utils._get_size(ndim, size)


@pytest.mark.parametrize("err_type, size, origin", [
    (TypeError, [1], 1.0), (TypeError, [1], [1.0]), (RuntimeError, [1], [[1]]),
    (RuntimeError, [1], [1, 1]), (ValueError, [1], [2]),
])
def test_errs__get_origin(err_type, size, origin):
    with pytest.raises(err_type):
        _utils._get_origin(size, origin)


@pytest.mark.parametrize("err_type, ndim, size, footprint", [
    (RuntimeError, 1, None, None), (RuntimeError, 1, [2], np.ones((2,), dtype=bool)),
    (RuntimeError, 1, None, np.ones((1, 2), dtype=bool)), (RuntimeError, 1, None, np.ones([0], dtype=bool)),
])
def test_errs__get_footprint(err_type, ndim, size, footprint):
    with pytest.raises(err_type):
        _utils._get_footprint(ndim, size=size, footprint=footprint)


@pytest.mark.parametrize("expected, ndim, depth, boundary", [
    (({0: 0}, {0: "none"}), 1, 0, "none"), (({0: 0}, {0: "reflect"}), 1, 0, "reflect"),
    (({0: 0}, {0: "periodic"}), 1, 0, "periodic"), (({0: 1}, {0: "none"}), 1, 1, "none"),
])
def test__get_depth_boundary(expected, ndim, depth, boundary):
    assert expected == _utils._get_depth_boundary(ndim, depth, boundary)


@pytest.mark.parametrize("expected, ndim, size", [((1,), 1, 1), ((3, 3), 2, 3), ((2, 4), 2, (2, 4))])
def test__get_size(expected, ndim, size):
    assert expected == _utils._get_size(ndim, size)


@pytest.mark.parametrize("expected, size, origin", [((0,), (1,), 0), ((1,), (3,), 1), ((1, 2), (3, 5), (1, 2))])
def test__get_origin(expected, size, origin):
    assert expected == _utils._get_origin(size, origin)


@pytest.mark.parametrize("expected, size, origin", [
    ((0,), (1,), 0), ((1,), (3,), 0), ((2,), (3,), 1), ((2, 4), (3, 5), (1, 2))])
def test__get_depth(expected, size, origin):
    assert expected == _utils._get_depth(size, origin)


@pytest.mark.parametrize("expected, ndim, size, footprint", [
    (np.ones((2,), dtype=bool), 1, 2, None), (np.ones((2,), dtype=bool), 1, None, np.ones((2,), dtype=bool))])
def test__get_footprint(expected, ndim, size, footprint):
    assert (expected == _utils._get_footprint(ndim, size, footprint)).all()


class _Shape2D:
    ndim = 2
    shape = (10, 14)
    dtype = np.dtype(float)


@pytest.mark.parametrize("err_type, sigma, truncate", [
    (RuntimeError, [[1.0]], 4.0), (RuntimeError, [1.0], 4.0), (TypeError, 1.0 + 0.0j, 4.0),
    (TypeError, 1.0, 4.0 + 0.0j),
])
@pytest.mark.parametrize("name", ["gaussian", "gaussian_filter", "gaussian_gradient_magnitude",
                                  "gaussian_laplace"])
def test_gaussian_filters_params(name, err_type, sigma, truncate):
    # reference: test__gaussian.py:12-35 -- raised before any dispatch / compute
    with pytest.raises(err_type):
        getattr(dask_image_b200.ndfilters, name)(_Shape2D(), sigma, truncate=truncate)


def test_gaussian_halo_is_ceil_sigma_truncate():
    assert _gaussian._get_border(_Shape2D(), (1.0, 2.3), 4.0) == (4, 10)
    sig = _gaussian._get_sigmas(_Shape2D(), 2)
    assert len(sig) == 2 and all(s == 2 for s in sig)


@pytest.mark.parametrize("err_type, size, origin", [
    (TypeError, 3.0, 0), (TypeError, 3, 0.0), (RuntimeError, [3], 0), (RuntimeError, 3, [0]),
    (RuntimeError, [[3]], 0), (RuntimeError, 3, [[0]]),
])
def test_uniform_filter_params(err_type, size, origin):
    # reference: test__smooth.py:12-28
    with pytest.raises(err_type):
        dask_image_b200.ndfilters.uniform_filter(_Shape2D(), size, origin=origin)


# ---- dispatch ----------------------------------------------------------------
def test_dispatcher_semantics():
    d = Dispatcher(name="demo")

    class A:
        pass

    class B(A):
        pass

    class Meta:
        def __init__(self, m):
            self._meta = m

    @d.register(A)
    def factory_a(*args, **kwargs):
        return "chunk-func-for-A"

    assert d(A()) == "chunk-func-for-A"
    assert d(B()) == "chunk-func-for-A"          # MRO walk
    assert d(Meta(B())) == "chunk-func-for-A"    # dask arrays dispatch on type(_meta)
    assert get_type(Meta(A())) is A
    with pytest.raises(TypeError):
        d(3.0)
    check_arraytypes_compatible(A(), Meta(A()))
    with pytest.raises(ValueError, match="Array types must be compatible"):
        check_arraytypes_compatible(A(), np.ones(3))


def test_b200array_is_registered_and_numpy_is_not():
    from dask_image_b200 import B200Array
    from dask_image_b200.dispatch import _dispatch_ndfilters as t

    for name in ["convolve", "correlate", "gaussian_filter", "gaussian_gradient_magnitude",
                 "gaussian_laplace", "uniform_filter", "laplace", "prewitt", "sobel"]:
        disp = getattr(t, "dispatch_" + name)
        assert disp.dispatch(B200Array)() is getattr(ndimage, name)
        with pytest.raises(TypeError):      # no CPU fallback registered
            disp.dispatch(np.ndarray)


def test_conv_params_host_side():
    # reference: test__conv.py:19-43 (conditions that do not need device data)
    nf = dask_image_b200.ndfilters
    img = _Shape2D()
    for f in (nf.convolve, nf.correlate):
        with pytest.raises(ValueError):   # array types differ
            f(img, np.ones((1, 1)))


# ---------------------------------------------------------------------------
# planning layer (pure host logic shared by the array, slab and streaming engines)
# ---------------------------------------------------------------------------
def test_plans_follow_scipy_pass_structure():
    from dask_image_b200 import _lib, ndimage as nd

    p = nd._plan_gaussian_filter(3, (2.0, 0.0, 1.0), order=(0, 0, 1), mode=("reflect", "wrap", "mirror"))
    (passes, op), = p.chains
    assert [ps.axis for ps in passes] == [0, 2] and op == _lib.EPI_STORE      # sigma 0 axis skipped
    assert [ps.mode for ps in passes] == ["reflect", "mirror"]
    assert len(passes[0].weights) == 2 * int(4 * 2.0 + 0.5) + 1
    assert nd._plan_gaussian_filter(2, 0.0).chains == []                       # identity
    g = nd._plan_gaussian_gradient_magnitude(3, 1.5)
    assert [op for _, op in g.chains] == [_lib.EPI_SQ_STORE, _lib.EPI_SQ_ACC, _lib.EPI_SQ_ACC_SQRT]
    assert all([ps.axis for ps in ch] == [0, 1, 2] for ch, _ in g.chains)
    g1 = nd._plan_gaussian_gradient_magnitude(1, 1.5)
    assert g1.zero_out and g1.chains[0][1] == _lib.EPI_SQ_ACC_SQRT
    lap = nd._plan_gaussian_laplace(2, (1.0, 2.0))
    assert [op for _, op in lap.chains] == [_lib.EPI_STORE, _lib.EPI_ACC]
    u = nd._plan_uniform_filter(3, (5, 1, 4), origin=(1, 0, -2))
    assert [(ps.kind, ps.axis, ps.size, ps.origin) for ps in u.chains[0][0]] == [("uniform", 0, 5, 1), ("uniform", 2, 4, -2)]
    with pytest.raises(ValueError):
        nd._plan_uniform_filter(2, 4, origin=2)
    s = nd._plan_sobel(3, axis=-1)
    assert [ps.axis for ps in s.chains[0][0]] == [2, 0, 1]                     # derivative axis first
    c = nd._plan_correlate_nd(2, np.dtype("f4"), np.arange(6.0).reshape(2, 3), origin=(0, 1), convolution=True)
    w, origins, mode = c.nd
    assert np.array_equal(w, np.arange(6.0).reshape(2, 3)[::-1, ::-1]) and origins == [-1, -1]
    m = nd._plan_minimum_filter(2, footprint=np.ones((3, 3), bool))
    assert [ps.kind for ps in m.chains[0][0]] == ["min", "min"] and m.extreme is None   # all-true footprint: separable
    m = nd._plan_maximum_filter(2, footprint=np.array([[0, 1, 0], [1, 1, 1], [0, 1, 0]], bool))
    assert m.extreme is not None and m.extreme[3] is True and m.chains == []
    with pytest.raises(RuntimeError):
        nd._plan_minimum_filter(2)


def test_chain_targets_and_prefix_signatures():
    from dask_image_b200 import _lib, ndimage as nd

    # store epilogues ping-pong between the output and ONE scratch volume
    assert nd._chain_targets(3, _lib.EPI_STORE) == ["out", 0, "out"]
    assert nd._chain_targets(2, _lib.EPI_SQ_STORE) == [0, "out"]
    assert nd._chain_targets(1, _lib.EPI_STORE) == ["out"]
    # accumulating epilogues must leave the output alone until the last pass
    assert nd._chain_targets(3, _lib.EPI_ACC) == [0, 1, "out"]
    assert nd._chain_targets(3, _lib.EPI_SQ_ACC_SQRT) == [0, 1, "out"]
    # the chains of ggm / laplace along axes 1 and 2 start with the same pass
    g = nd._plan_gaussian_gradient_magnitude(3, 2.0)
    sig = [tuple(nd._pass_signature(ps) for ps in ch) for ch, _ in g.chains]
    assert sig[1][0] == sig[2][0] and sig[0][0] != sig[1][0] and sig[1][1] != sig[2][1]


def test_slab_halo_eligibility_is_a_pure_host_query():
    """b200_correlate1d_halo_eligible needs no GPU and depends on the geometry
    only, so every rank can evaluate it for every slab size."""
    from dask_image_b200 import ndimage as nd

    f4 = np.dtype("f4")
    assert nd._halo_eligible(f4, 256, 2048 * 2048, 17, 0)
    assert nd._halo_eligible(np.dtype("f8"), 300, 512, 9, 2)
    assert not nd._halo_eligible(f4, 20, 2048 * 2048, 17, 0)          # thinner than one tile + halo
    assert not nd._halo_eligible(f4, 256, 2047, 17, 0)                # rows not 16-byte multiples
    assert not nd._halo_eligible(np.dtype("i4"), 256, 2048, 17, 0)
    u2 = np.dtype("u2")
    assert nd._uniform_halo_eligible(u2, 256, 4096, 7, 0, "reflect", 0.0)
    assert not nd._uniform_halo_eligible(u2, 256, 4096, 7, 0, "constant", 0.5)    # cval is not a uint16
    assert not nd._uniform_halo_eligible(u2, 256, 4096, 200, 0, "reflect", 0.0)   # window too long for exact fp32 sums
    assert nd._uniform_halo_eligible(f4, 256, 4096, 7, 0, "constant", 0.5)


# ---- remaining parameter tables of the reference's tests ------------------------
@pytest.mark.parametrize("err_type, axis", [(ValueError, 0.0), (ValueError, 2), (ValueError, -3)])
@pytest.mark.parametrize("name", ["prewitt", "sobel"])
def test_edge_func_params(name, err_type, axis):
    # reference: test__edge.py:15-37
    with pytest.raises(err_type):
        getattr(dask_image_b200.ndfilters, name)(_Shape2D(), axis)


@pytest.mark.parametrize("err_type, size, footprint, origin", [
    (RuntimeError, None, None, 0), (TypeError, 1.0, None, 0), (RuntimeError, (1,), None, 0),
    (RuntimeError, [(1,)], None, 0), (RuntimeError, 1, np.ones((1,)), 0), (RuntimeError, None, np.ones((1,)), 0),
    (RuntimeError, None, np.ones((1, 0)), 0), (RuntimeError, 1, None, (0,)), (RuntimeError, 1, None, [(0,)]),
    (ValueError, 1, None, 1), (TypeError, 1, None, 0.0), (TypeError, 1, None, (0.0, 0.0)),
    (TypeError, 1, None, 1 + 0j), (TypeError, 1, None, (0 + 0j, 1 + 0j)),
])
@pytest.mark.parametrize("name", ["minimum_filter", "maximum_filter"])
def test_order_filter_params(name, err_type, size, footprint, origin):
    # reference: test__order.py:20-66 -- all raised before any dispatch / compute
    with pytest.raises(err_type):
        getattr(dask_image_b200.ndfilters, name)(_Shape2D(), size=size, footprint=footprint, origin=origin)


def test_out_of_scope_entry_points_raise_one_clear_error():
    """dask_image.ndfilters names outside the correlation path exist and raise OutOfScopeError
    (a NotImplementedError) -- never an AttributeError, never a silent CPU fallback."""
    from dask_image_b200 import ndfilters

    reference_all = {"convolve", "correlate", "laplace", "prewitt", "sobel", "gaussian", "gaussian_filter",
                     "gaussian_gradient_magnitude", "gaussian_laplace", "generic_filter", "minimum_filter",
                     "median_filter", "maximum_filter", "rank_filter", "percentile_filter", "uniform_filter",
                     "threshold_local"}                      # ref:dask_image/ndfilters/__init__.py:3-21
    assert reference_all <= set(ndfilters.__all__)
    for name in ("generic_filter", "median_filter", "rank_filter", "percentile_filter"):
        f = getattr(ndfilters, name)
        assert f.__name__ == name and f.__module__ == "dask_image_b200.ndfilters"
        with pytest.raises(ndfilters.OutOfScopeError, match=name):
            f(object(), size=3)
    assert issubclass(ndfilters.OutOfScopeError, NotImplementedError)
