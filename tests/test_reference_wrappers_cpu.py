"""The UNMODIFIED reference wrappers driving this package's chunk functions.

dask is not installable here, but the only thing the reference's ndfilters path imports from it is
``dask.utils.Dispatch`` (ref:dask_image/dispatch/_dispatcher.py:3); everything else it needs from a dask array are
the attributes ``ndim`` / ``dtype`` / ``_meta`` and the method ``map_overlap`` (ref:dask_image/ndfilters/_gaussian.py:70-81).
This test puts a stand-in for that one class in ``sys.modules`` (tests/_reference_stub.py), imports
``dask_image.ndfilters`` from /root/reference (or from its pip-installed copy in baseline/_ref) as it is, lets
``register_with_dask_image()`` add the ``B200Array`` registrations to the REFERENCE's dispatch tables (the lines a
maintainer would add) and calls the reference's public functions on ``HaloChunked`` arrays
(the eager ``map_overlap`` stand-in).  The launches below the C ABI are replaced by the oracle as in
test_host_path_cpu.py, so what is proven is the drop-in contract: the reference's validation, halo depths, dispatch
and the exact kwargs it forwards per chunk (sigma as a tuple of numpy scalars, the footprint array, the origin as
given, B200Array weights ...) reach chunk functions that accept them, and the result is whole-array scipy, bit for
bit.  It mirrors the reference's cupy smoke tests (ref:tests/test_dask_image/test_ndfilters/test_cupy_ndfilters.py:13-103)
with value checks.  tests/test_gpu_reference_wrappers.py repeats it on the GPU with the real kernels below the C ABI;
tests/test_gpu_dask.py is the same thing under a real dask.
"""
import numpy as np
import pytest
import scipy.ndimage as ndi

import _reference_stub
from test_host_path_cpu import _data, cpu_path  # noqa: F401  (fixture)

pytestmark = pytest.mark.skipif(_reference_stub.unavailable_reason() is not None,
                                reason=str(_reference_stub.unavailable_reason()))


@pytest.fixture
def ref_ndfilters(cpu_path):  # noqa: F811
    import dask_image_b200

    with _reference_stub.reference_ndfilters() as ref:
        yield ref, dask_image_b200


def _eq(result, ref):
    got = np.asarray(result.compute())
    assert got.dtype == ref.dtype and got.shape == ref.shape
    np.testing.assert_array_equal(got, ref)


@pytest.mark.parametrize("dtype", [np.float32, np.uint16])
@pytest.mark.parametrize("chunks", [(5, 5), (4, 7), (10, 12)])
def test_reference_wrappers_route_b200_chunks(ref_ndfilters, dtype, chunks):
    ref, b200 = ref_ndfilters
    from dask_image.dispatch import _dispatch_ndfilters as ref_tables

    a = _data(np.random.default_rng(2), (10, 12), dtype)
    d = b200.from_array(a, chunks=chunks)
    # the factory the REFERENCE's table hands out for this array is this package's chunk function
    assert ref_tables.dispatch_gaussian_filter(d) is b200.ndimage.gaussian_filter
    assert ref_tables.dispatch_convolve(d) is b200.ndimage.convolve
    with np.errstate(all="ignore"):
        for name in ("gaussian", "gaussian_filter"):
            _eq(getattr(ref, name)(d, 1), ndi.gaussian_filter(a, 1))
            _eq(getattr(ref, name)(d, (1.0, 0.5), order=(0, 1), mode="mirror", truncate=3.0),
                ndi.gaussian_filter(a, (1.0, 0.5), order=(0, 1), mode="mirror", truncate=3.0))
        _eq(ref.gaussian_gradient_magnitude(d, 1), ndi.gaussian_gradient_magnitude(a, 1))
        _eq(ref.gaussian_laplace(d, 1, mode="nearest"), ndi.gaussian_laplace(a, 1, mode="nearest"))
        _eq(ref.uniform_filter(d, 3), ndi.uniform_filter(a, 3))
        _eq(ref.uniform_filter(d, (2, 5), origin=(-1, 1), mode="constant", cval=2.0),
            ndi.uniform_filter(a, (2, 5), origin=(-1, 1), mode="constant", cval=2.0))
        _eq(ref.laplace(d), ndi.laplace(a))
        _eq(ref.prewitt(d), ndi.prewitt(a))
        _eq(ref.sobel(d, axis=0, mode="mirror"), ndi.sobel(a, axis=0, mode="mirror"))
        w = np.random.default_rng(3).random((3, 3)) - 0.2
        wd = b200.B200Array.from_numpy(w)         # ref:dask_image/dispatch/_utils.py:23-25: same array type
        for name in ("convolve", "correlate"):
            _eq(getattr(ref, name)(d, wd), getattr(ndi, name)(a, w))
            # mode="wrap" becomes a periodic halo + mode="constant" per chunk (ref:_conv.py:23-25,50-52)
            _eq(getattr(ref, name)(d, wd, mode="wrap", origin=(1, -1)), getattr(ndi, name)(a, w, mode="wrap", origin=(1, -1)))
            with pytest.raises(ValueError, match="Array types must be compatible"):
                getattr(ref, name)(d, w)
        # ref:tests/.../test_cupy_ndfilters.py:64-103: size / footprint forms of the order filters
        for name in ("minimum_filter", "maximum_filter"):
            for size, fp in [(1, None), ((1, 1), None), (None, np.ones((1, 1))), (3, None), ((3, 2), None),
                             (None, np.array([[0, 1, 0], [1, 1, 1], [0, 1, 0]]))]:
                _eq(getattr(ref, name)(d, size=size, footprint=fp), getattr(ndi, name)(a, size=size, footprint=fp))


def test_reference_threshold_local_gaussian(ref_ndfilters):
    ref, b200 = ref_ndfilters
    a = _data(np.random.default_rng(4), (10, 12), np.uint8)
    d = b200.from_array(a, chunks=(5, 6))
    got = np.asarray(ref.threshold_local(d, 7, offset=2.5, mode="mirror").compute())
    want = ndi.gaussian_filter(a.astype(np.float64), (7 - 1) / 6.0, mode="mirror") - 2.5
    np.testing.assert_array_equal(got, want)


def test_reference_validation_runs_before_any_launch(ref_ndfilters, cpu_path):  # noqa: F811
    """The reference's own argument checks (ref:tests/.../test__gaussian.py:12-35, test__smooth.py:12-28,
    test__conv.py:19-43) fire unchanged for B200Array-backed arrays, before anything is launched."""
    ref, b200 = ref_ndfilters
    d = b200.from_array(np.zeros((6, 6), np.float32), chunks=(3, 3))
    del cpu_path[:]
    with pytest.raises(RuntimeError):
        ref.gaussian_filter(d, (1.0, 1.0, 1.0))
    with pytest.raises(TypeError):
        ref.gaussian_filter(d, "x")
    with pytest.raises(TypeError):
        ref.uniform_filter(d, 3.0)
    with pytest.raises(ValueError):
        ref.uniform_filter(d, 3, origin=2)
    with pytest.raises(ValueError):
        ref.convolve(d, b200.B200Array.from_numpy(np.ones((1,))))
    with pytest.raises(RuntimeError):
        ref.convolve(d, b200.B200Array.from_numpy(np.ones((1, 1))), origin=(0,))
    assert cpu_path == []
    # out of scope: no B200Array registration, the reference reports it as a dispatch TypeError
    with pytest.raises(TypeError):
        ref.median_filter(d, size=3)


# --------------------------------------------------------------------------
# the package's own wrappers against the reference's, and against the reference's numpy path
# --------------------------------------------------------------------------
class NumpyChunked:
    """numpy twin of HaloChunked (test infrastructure): the surface the reference touches on a dask array, with
    ``map_overlap(depth, boundary in {"none", "periodic"})`` evaluated eagerly chunk by chunk.  Its chunks are
    numpy arrays, so the reference's OWN registrations (scipy.ndimage) serve it: this is the reference's CPU path."""

    def __init__(self, a, chunks):
        self._a = np.ascontiguousarray(a)
        self.chunksize = tuple(chunks)
        self._meta = np.empty((0,) * self._a.ndim, dtype=self._a.dtype)

    ndim = property(lambda self: self._a.ndim)
    shape = property(lambda self: self._a.shape)
    dtype = property(lambda self: self._a.dtype)

    def compute(self):
        return self._a

    def astype(self, dtype):
        return NumpyChunked(self._a.astype(dtype), self.chunksize)

    def __sub__(self, other):
        return NumpyChunked(self._a - other, self.chunksize)

    def map_overlap(self, func, depth, boundary, dtype=None, meta=None, **kwargs):
        import itertools

        nd = self.ndim
        # dask accepts a scalar or a per-axis mapping / sequence for both arguments
        depth = [int(depth[d]) if hasattr(depth, "__getitem__") else int(depth) for d in range(nd)]
        boundary = [boundary if (boundary is None or isinstance(boundary, str)) else boundary[d] for d in range(nd)]
        src = self._a
        shift = [0] * nd
        for d in range(nd):
            if boundary[d] == "periodic" and depth[d] > 0:
                src = np.take(src, np.arange(-depth[d], src.shape[d] + depth[d]) % src.shape[d], axis=d)
                shift[d] = depth[d]
            else:
                assert boundary[d] in ("none", None, "periodic")
        out = np.empty_like(self._a)
        spans = [[(s, min(s + self.chunksize[d], self.shape[d])) for s in range(0, self.shape[d], self.chunksize[d])]
                 for d in range(nd)]
        for job in itertools.product(*spans):
            src_sl, trim_sl, dst_sl = [], [], []
            for d, (s, e) in enumerate(job):
                lo, hi = max(s + shift[d] - depth[d], 0), min(e + shift[d] + depth[d], src.shape[d])
                src_sl.append(slice(lo, hi))
                trim_sl.append(slice(s + shift[d] - lo, s + shift[d] - lo + e - s))
                dst_sl.append(slice(s, e))
            res = func(np.ascontiguousarray(src[tuple(src_sl)]), **kwargs)
            out[tuple(dst_sl)] = res[tuple(trim_sl)]
        return NumpyChunked(out, self.chunksize)


CALLS = [
    ("gaussian_filter", dict(sigma=1.3)),
    ("gaussian_filter", dict(sigma=(0.0, 2.0), order=(0, 2), truncate=2.0)),
    ("gaussian_gradient_magnitude", dict(sigma=0.9)),
    ("gaussian_laplace", dict(sigma=(1.1, 0.7), truncate=3.0)),
    ("uniform_filter", dict(size=(4, 3), origin=(1, -1))),
    ("laplace", dict()),
    ("prewitt", dict(axis=0)),
    ("sobel", dict()),
    ("convolve", dict(weights=np.array([[0.2, -1.0, 0.5], [0.1, 0.3, -0.4]]), origin=(0, 1))),
    ("correlate", dict(weights=np.arange(12.0).reshape(4, 3) / 7 - 0.5)),
    ("minimum_filter", dict(size=(3, 2), origin=(1, 0))),
    ("maximum_filter", dict(footprint=np.array([[1, 0, 1], [0, 1, 0]], bool))),
]


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.uint8, np.int16])
@pytest.mark.parametrize("mode", ["reflect", "constant", "nearest", "mirror", "wrap"])
def test_three_ways_agree(ref_ndfilters, dtype, mode):
    """(i) the reference on numpy chunks = the reference's CPU path (its wrappers, its scipy registrations,
    ``map_overlap`` emulated), (ii) the reference's wrappers on B200Array chunks, (iii) this package's mirror of the
    wrappers on B200Array chunks: bit-identical, for every boundary mode -- including the reference's own quirk that
    gaussian / uniform / order filters wrap INSIDE each edge chunk for mode='wrap' (boundary stays "none",
    ref:_gaussian.py:70-81) while convolve / correlate wrap around the whole array (ref:_conv.py:23-25)."""
    ref, b200 = ref_ndfilters
    a = _data(np.random.default_rng(8), (13, 11), dtype)
    for chunks in [(5, 4), (13, 11), (7, 11)]:
        dn, dd = NumpyChunked(a, chunks), b200.from_array(a, chunks=chunks)
        for name, kw in CALLS:
            kw = dict(kw, mode=mode, cval=1.0)
            kwd = dict(kw)
            if "weights" in kw:
                kwd["weights"] = b200.B200Array.from_numpy(kw["weights"])
            with np.errstate(all="ignore"):
                want = getattr(ref, name)(dn, **kw).compute()
                via_ref = np.asarray(getattr(ref, name)(dd, **kwd).compute())
                own = np.asarray(getattr(b200.ndfilters, name)(dd, **kwd).compute())
            np.testing.assert_array_equal(via_ref, want, err_msg="reference wrappers, %s %s %s" % (name, kw, chunks))
            np.testing.assert_array_equal(own, want, err_msg="own wrappers, %s %s %s" % (name, kw, chunks))
            if chunks == a.shape or mode != "wrap" or name in ("convolve", "correlate"):
                # and that is whole-array scipy (map_overlap's invariant)
                whole = getattr(ndi, name)(a, **kw)
                if name == "uniform_filter" and a.dtype.kind == "f":
                    # scipy's float box filter is a running sum along each line: its rounding depends on where
                    # the line starts, so the REFERENCE itself is chunking-invariant only to rounding here
                    np.testing.assert_allclose(want, whole, rtol=1e-5 if a.dtype == np.float32 else 1e-12)
                else:
                    np.testing.assert_array_equal(want, whole)


def test_goldens_are_the_reference_outputs(ref_ndfilters):
    """Pins the committed fixtures (tests/golden/, generated with scipy by make_golden.py) to outputs of the
    REFERENCE ITSELF run here: every golden case of a public dask_image.ndfilters function is recomputed through the
    reference's wrappers and numpy registrations -- as one chunk, and cut into chunks wherever chunking is exact
    (everything but the float running-sum box filter and the wrap-inside-a-chunk quirk) -- and must equal the stored
    output bit for bit."""
    from _helpers import golden_case, golden_cases

    ref, _ = ref_ndfilters
    public = {"gaussian_filter", "gaussian_gradient_magnitude", "gaussian_laplace", "uniform_filter", "convolve",
              "correlate"}
    checked = chunked = 0
    for i, spec in enumerate(golden_cases()):
        if spec["func"] not in public:
            continue
        func, a, kw, want = golden_case(i)
        with np.errstate(all="ignore"):
            got = getattr(ref, func)(NumpyChunked(a, a.shape), **kw).compute()
            np.testing.assert_array_equal(got, want, err_msg=spec["name"])
            checked += 1
            wraps_in_chunk = kw.get("mode") == "wrap" and func not in ("convolve", "correlate")
            if not wraps_in_chunk and not (func == "uniform_filter" and a.dtype.kind == "f"):
                chunks = tuple(max(1, (n + 2) // 3) for n in a.shape)
                got = getattr(ref, func)(NumpyChunked(a, chunks), **kw).compute()
                np.testing.assert_array_equal(got, want, err_msg=spec["name"] + " chunked")
                chunked += 1
    assert checked > 300 and chunked > 250, (checked, chunked)


def test_integration_md_registration_snippet_works(cpu_path):  # noqa: F811
    """The patch INTEGRATION.md section 1 proposes to a dask-image maintainer, taken from the document verbatim and
    executed inside the reference's ``dask_image/dispatch/_dispatch_ndfilters.py`` namespace: lazy registration
    keyed on the ``dask_image_b200`` top-level module, fired by the first B200Array chunk."""
    import os
    import re
    import sys
    import types

    import dask_image_b200 as b200
    from dask_image_b200.dispatch._dispatcher import _Base

    text = open(os.path.join(_reference_stub.ROOT, "INTEGRATION.md")).read()
    section = text[text.index("## 1. The registration a maintainer adds"):text.index("## 2.")]
    snippet = re.search(r"```python\n(.*?)```", section, re.S).group(1)
    assert "register_lazy(\"dask_image_b200\"" in snippet

    root = _reference_stub.reference_root()
    dask, utils = types.ModuleType("dask"), types.ModuleType("dask.utils")
    utils.Dispatch = _Base
    dask.utils = utils
    saved = {k: sys.modules.get(k) for k in ("dask", "dask.utils")}
    sys.modules["dask"], sys.modules["dask.utils"] = dask, utils
    sys.path.insert(0, root)
    before = set(sys.modules)
    try:
        import dask_image.ndfilters as ref
        from dask_image.dispatch import _dispatch_ndfilters as tables

        d = b200.from_array(np.arange(48, dtype=np.float32).reshape(6, 8), chunks=(3, 4))
        with pytest.raises(TypeError):                   # nothing registered for B200Array yet
            tables.dispatch_gaussian_filter(d)
        exec(compile(snippet, "INTEGRATION.md#1", "exec"), tables.__dict__)
        got = np.asarray(ref.gaussian_filter(d, 1.0).compute())
        np.testing.assert_array_equal(got, ndi.gaussian_filter(np.arange(48, dtype=np.float32).reshape(6, 8), 1.0))
        assert tables.dispatch_uniform_filter(d) is b200.ndimage.uniform_filter
        assert tables.dispatch_maximum_filter(d) is b200.ndimage.maximum_filter
    finally:
        sys.path.remove(root)
        for name in set(sys.modules) - before:
            if name == "dask_image" or name.startswith("dask_image."):
                del sys.modules[name]
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
