"""GPU tests of the TMA-tiled kernels (corr1d_strided.cu, corr1d_contig.cu, corr1d_minmax.cu): forced onto the tiled
path with B200_FLAG_FORCE_FAST so that small, edge-heavy shapes exercise tile
edges, patched halos, tap tails and both axis orientations; compared against
the CPU oracle within the north_star tolerances (fp32 1e-5, fp64 1e-12)."""
import contextlib

import numpy as np
import pytest
import scipy.ndimage as ndi

from _helpers import assert_parity
from oracle import oracle as orc

pytestmark = pytest.mark.gpu

MODES = ["reflect", "wrap", "nearest", "constant", "mirror"]


@pytest.fixture(scope="module")
def b200():
    import dask_image_b200

    return dask_image_b200


@contextlib.contextmanager
def forced(b200, flag):
    old = b200.ndimage.set_default_flags(flag)
    try:
        yield
    finally:
        b200.ndimage.set_default_flags(old)


def dev(b200, a):
    return b200.B200Array.from_numpy(np.asarray(a))


SHAPES_AXES = [
    ((70, 64), 0), ((70, 64), 1), ((5, 130, 36), 1), ((3, 40, 132), 2), ((200, 8, 16), 0),
    ((260, 512), 0), ((260, 512), 1), ((9, 260), 1), ((2, 3, 24, 20), 2), ((1, 4), 1), ((3, 8), 0),
]
TAPS = [(1, 0), (2, -1), (3, 1), (8, 0), (9, 0), (17, 0), (19, 0), (24, 3), (25, -4), (33, 5), (41, 0), (64, -32),
        (101, 7)]


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("shape, axis", SHAPES_AXES)
def test_tma_correlate1d(b200, dtype, mode, shape, axis):
    from dask_image_b200 import _lib

    rng = np.random.default_rng(hash((shape, axis, mode)) % (2 ** 32))
    a = (rng.random(shape) - 0.25).astype(dtype)
    d = dev(b200, a)
    for nw, origin in TAPS:
        w = rng.random(nw) - 0.3
        with forced(b200, _lib.FLAG_FORCE_FAST):
            got = b200.ndimage.correlate1d(d, w, axis=axis, mode=mode, cval=0.75, origin=origin)
        assert _lib.last_kernel().startswith("corr1d_tma_"), _lib.last_kernel()
        ref = orc.correlate1d(a, w, axis=axis, mode=mode, cval=0.75, origin=origin)
        assert_parity(got.to_numpy(), ref, what="nw=%d origin=%d" % (nw, origin))


@pytest.mark.parametrize("mode", MODES)
def test_tma_constant_cval_zero_uses_tma_zero_fill(b200, mode):
    from dask_image_b200 import _lib

    a = np.random.default_rng(0).random((150, 256)).astype(np.float32)
    w = np.random.default_rng(1).random(21)
    for axis in (0, 1):
        with forced(b200, _lib.FLAG_FORCE_FAST):
            got = b200.ndimage.correlate1d(dev(b200, a), w, axis=axis, mode=mode, cval=0.0)
        assert_parity(got.to_numpy(), orc.correlate1d(a, w, axis=axis, mode=mode, cval=0.0))


UNALIGNED_WIDTHS = list(range(1, 68)) + [110, 1023, 1034]


@pytest.mark.parametrize("width", UNALIGNED_WIDTHS)
def test_unaligned_rows_stay_on_the_tiled_kernels(b200, width):
    """Rows that are not 16-byte multiples (dask chunks with odd halos: ceil(4 sigma) = 5, 7, 9 ...) cannot be
    TMA-staged; the same tiles are then filled with plain loads and stored element-wise -- the arithmetic is
    that of the tiled kernels, never the one-thread-per-output kernel (a ~20x cliff in round 1)."""
    from dask_image_b200 import _lib

    rng = np.random.default_rng(width)
    a = (rng.random((37, width)) - 0.25).astype(np.float32)
    d = dev(b200, a)
    aligned = width % 4 == 0
    for nw, origin, mode in ((5, 0, "reflect"), (17, 0, "mirror"), (9, -2, "wrap"), (12, 3, "constant"), (25, 0, "nearest")):
        w = rng.random(nw) - 0.3
        for axis in (0, 1):
            with forced(b200, _lib.FLAG_FORCE_FAST):
                got = b200.ndimage.correlate1d(d, w, axis=axis, mode=mode, cval=0.7, origin=origin)
            k = _lib.last_kernel()
            if width == 1 and axis == 0:
                want = "corr1d_ldg_contig_f32"     # a single column IS a contiguous line (37 * 4 bytes: unaligned)
            else:
                want = "corr1d_%s_%s_f32" % ("tma" if aligned else "ldg", "strided" if axis == 0 else "contig")
            assert k == want, (width, axis, k)
            assert_parity(got.to_numpy(), orc.correlate1d(a, w, axis=axis, mode=mode, cval=0.7, origin=origin),
                          what="width=%d nw=%d axis=%d" % (width, nw, axis))
    # uint16 / uint8: the exact box filter, and the weighted integer kernels (plain loads by construction)
    for dt in (np.uint16, np.uint8):
        u = rng.integers(0, np.iinfo(dt).max, (21, 19, width), dtype=dt, endpoint=True)
        du = dev(b200, u)
        got = b200.ndimage.uniform_filter(du, 5, mode="mirror")
        assert "generic" not in _lib.last_kernel(), (width, _lib.last_kernel())
        assert np.array_equal(got.to_numpy(), orc.uniform_filter(u, 5, mode="mirror"))
        got = b200.ndimage.gaussian_filter(du, 1.0)
        assert _lib.last_kernel().startswith("corr1d_int_"), (width, _lib.last_kernel())
        assert np.array_equal(got.to_numpy(), orc.gaussian_filter(u, 1.0))
    # float32 volume: axis-0 pass + fused pair, whatever the width
    v = rng.random((9, 23, width)).astype(np.float32)
    _lib.launch_count(reset=True)
    got = b200.ndimage.gaussian_filter(dev(b200, v), 1.5, mode="reflect")
    assert _lib.launch_count() == 2 and _lib.last_kernel().startswith("corr2d_fused_"), _lib.last_kernel()
    assert_parity(got.to_numpy(), orc.gaussian_filter(v, 1.5, mode="reflect"))


def test_unaligned_base_pointer(b200):
    """A view that starts 4 bytes into an allocation: 16-byte rows, unaligned base -> plain-load staging."""
    import torch

    from dask_image_b200 import _lib, ndimage

    rng = np.random.default_rng(12)
    a = rng.random((40 * 64 + 1,)).astype(np.float32)
    t = torch.from_numpy(a).cuda()
    x = t[1:].view(40, 64)
    assert x.data_ptr() % 16 == 4 and x.is_contiguous()
    w = rng.random(7)
    for axis in (0, 1):
        out = torch.empty(40, 64, device="cuda")
        ndimage._launch_correlate1d(x, out, (40, 64), np.dtype(np.float32), axis, w, "reflect", 0.0, 0)
        assert _lib.last_kernel() == "corr1d_ldg_%s_f32" % ("strided" if axis == 0 else "contig")
        assert_parity(out.cpu().numpy(), orc.correlate1d(a[1:].reshape(40, 64), w, axis=axis))


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("name", ["gaussian_filter", "gaussian_gradient_magnitude", "gaussian_laplace"])
def test_gaussian_family_auto_selects_tiled_kernels(b200, dtype, name):
    from dask_image_b200 import _lib

    a = np.random.default_rng(3).random((48, 80, 96)).astype(dtype)
    got = getattr(b200.ndimage, name)(dev(b200, a), 2.0)
    # float32: the last two passes run as one fused launch (corr2d_fused.cu)
    want = "corr2d_fused_tma" if dtype == np.float32 else "corr1d_tma_contig"
    assert _lib.last_kernel().startswith(want), _lib.last_kernel()
    assert_parity(got.to_numpy(), getattr(ndi, name)(a, 2.0), what=name)


@pytest.mark.parametrize("sigma", [2.0, 4.0, (4.0, 1.0, 7.5)])
def test_gaussian_volume_vs_scipy(b200, sigma):
    """A volume big enough for interior, edge and corner tiles on every axis."""
    a = np.random.default_rng(1234).random((160, 200, 264), dtype=np.float32)
    got = b200.ndimage.gaussian_filter(dev(b200, a), sigma)
    assert_parity(got.to_numpy(), ndi.gaussian_filter(a, sigma))


def test_config1_4096sq_through_dispatch_boundary(b200):
    """BASELINE.json configs[0]: gaussian_filter sigma=2 on 4096^2 float32 in
    1024^2 chunks, driven chunk-wise through the dispatch boundary."""
    a = np.random.default_rng(1234).random((4096, 4096), dtype=np.float32)
    ref = ndi.gaussian_filter(a, 2.0)
    d = b200.from_array(a, chunks=(1024, 1024))
    got = b200.ndfilters.gaussian_filter(d, sigma=2)
    assert_parity(np.asarray(got), ref)
    whole = b200.ndfilters.gaussian_filter(dev(b200, a), sigma=2)
    assert_parity(np.asarray(whole), ref)
    # chunked and whole-array runs agree bit-for-bit (direct-form correlation)
    assert np.array_equal(np.asarray(whole), np.asarray(got))


def test_launch_counter_and_kernel_names(b200):
    from dask_image_b200 import _lib

    a = np.random.default_rng(5).random((64, 64, 64), dtype=np.float32)
    _lib.launch_count(reset=True)
    b200.ndimage.gaussian_filter(dev(b200, a), 2.0)
    assert _lib.launch_count() == 2      # axis-0 pass + the fused pair of the last two axes
    _lib.launch_count(reset=True)
    b200.ndimage.gaussian_laplace(dev(b200, a), 2.0)
    assert _lib.launch_count() == 5      # scipy's 9 passes: 2 axis-0 passes (one shared) + 3 fused pairs
    old = b200.ndimage.set_pair_fusion(False)
    try:
        _lib.launch_count(reset=True)
        b200.ndimage.gaussian_filter(dev(b200, a), 2.0)
        assert _lib.launch_count() == 3
        _lib.launch_count(reset=True)
        b200.ndimage.gaussian_laplace(dev(b200, a), 2.0)
        assert _lib.launch_count() == 8  # the shared axis-0 smoothing runs once
    finally:
        b200.ndimage.set_pair_fusion(old)


# ---------------------------------------------------------------------------
# integer box filter on the tiled kernels (exact: window sums in fp32, division
# by the proven float epilogue) -- bit-identical to the oracle
# ---------------------------------------------------------------------------
U_SHAPES_AXES = [((70, 64), 0), ((70, 64), 1), ((5, 130, 48), 1), ((3, 40, 144), 2), ((200, 8, 16), 0),
                 ((260, 512), 0), ((9, 272), 1), ((2, 3, 24, 32), 2), ((300, 16), 0), ((4, 16), 1)]
U_SIZES = [(2, 0), (2, -1), (3, 0), (3, 1), (7, 0), (7, -3), (8, 3), (16, -8), (31, 4), (64, 0), (128, 63)]


@pytest.mark.parametrize("dtype", [np.uint16, np.uint8])
@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("shape, axis", U_SHAPES_AXES)
def test_tma_uniform_integer_exact(b200, dtype, mode, shape, axis):
    from dask_image_b200 import _lib

    rng = np.random.default_rng(hash((shape, axis, mode)) % (2 ** 32))
    hi = np.iinfo(dtype).max
    a = rng.integers(0, hi + 1, shape).astype(dtype)
    a.flat[:: 7] = hi          # saturated runs: the largest window sums
    a.flat[3:: 11] = 0
    d = dev(b200, a)
    for size, origin in U_SIZES:
        with forced(b200, _lib.FLAG_FORCE_FAST):
            got = b200.ndimage.uniform_filter1d(d, size, axis=axis, mode=mode, cval=float(hi // 3),
                                                origin=origin)
        assert _lib.last_kernel().startswith("uniform1d_tma_"), _lib.last_kernel()
        ref = orc.uniform_filter1d(a, size, axis=axis, mode=mode, cval=float(hi // 3), origin=origin)
        assert_parity(got.to_numpy(), ref, exact=True, what="size=%d origin=%d" % (size, origin))


def test_tma_uniform_all_saturated_and_all_quotients(b200):
    """Every window sum of a saturated volume, and a ramp that walks the
    quotient classes, divide exactly."""
    from dask_image_b200 import _lib

    for dtype in (np.uint16, np.uint8):
        hi = np.iinfo(dtype).max
        for a in (np.full((64, 256), hi, dtype), (np.arange(64 * 256) % (hi + 1)).astype(dtype).reshape(64, 256)):
            for size in (2, 3, 5, 7, 9, 11, 100, 127, 128):
                for axis in (0, 1):
                    got = b200.ndimage.uniform_filter1d(dev(b200, a), size, axis=axis, mode="nearest")
                    assert _lib.last_kernel().startswith("uniform1d_tma_"), _lib.last_kernel()
                    assert_parity(got.to_numpy(), ndi.uniform_filter1d(a, size, axis=axis, mode="nearest"),
                                  exact=True, what="%s size=%d axis=%d" % (np.dtype(dtype), size, axis))


def test_uniform_integer_falls_back_when_not_provable(b200):
    from dask_image_b200 import _lib

    a = np.random.default_rng(0).integers(0, 65536, (64, 64)).astype(np.uint16)
    d = dev(b200, a)
    # cval that is not a uint16 value, a window too long for exact fp32 sums, a signed dtype
    got = b200.ndimage.uniform_filter1d(d, 5, axis=0, mode="constant", cval=0.5)
    assert _lib.last_kernel() == "uniform1d_generic"
    assert_parity(got.to_numpy(), ndi.uniform_filter1d(a, 5, axis=0, mode="constant", cval=0.5), exact=True)
    got = b200.ndimage.uniform_filter1d(d, 200, axis=1)
    assert _lib.last_kernel() == "uniform1d_generic"
    assert_parity(got.to_numpy(), ndi.uniform_filter1d(a, 200, axis=1), exact=True)
    s = a.astype(np.int16)
    got = b200.ndimage.uniform_filter1d(dev(b200, s), 5, axis=1)
    assert _lib.last_kernel() == "uniform1d_generic"
    assert_parity(got.to_numpy(), ndi.uniform_filter1d(s, 5, axis=1), exact=True)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_uniform_float_uses_tiled_correlation(b200, dtype):
    from dask_image_b200 import _lib

    a = np.random.default_rng(4).random((48, 72, 96)).astype(dtype)
    got = b200.ndimage.uniform_filter(dev(b200, a), (7, 4, 9), mode="mirror", origin=(0, -1, 2))
    assert _lib.last_kernel().startswith("corr1d_tma_contig"), _lib.last_kernel()
    assert_parity(got.to_numpy(), ndi.uniform_filter(a, (7, 4, 9), mode="mirror", origin=(0, -1, 2)))


def test_config3_uniform_u16_volume(b200):
    """BASELINE.json configs[2] at a checkable size: uniform_filter size=7 on a
    uint16 volume, every pass on the tiled integer kernels, bit-exact."""
    from dask_image_b200 import _lib

    a = np.random.default_rng(1234).integers(0, 65536, (136, 200, 264), dtype=np.uint16)
    ref = ndi.uniform_filter(a, 7)
    _lib.launch_count(reset=True)
    got = b200.ndfilters.uniform_filter(dev(b200, a), size=7)
    # the axis-0 pass + the fused pair along the last two axes (corr2d_fused.cu)
    assert _lib.launch_count() == 2 and _lib.last_kernel() == "uniform2d_fused_tma_u16"
    assert_parity(got.to_numpy(), ref, exact=True)
    old = b200.ndimage.set_pair_fusion(False)
    try:
        _lib.launch_count(reset=True)
        got = b200.ndfilters.uniform_filter(dev(b200, a), size=7)
        assert _lib.launch_count() == 3 and _lib.last_kernel() == "uniform1d_tma_contig_u16"
        assert_parity(got.to_numpy(), ref, exact=True)
    finally:
        b200.ndimage.set_pair_fusion(old)


# ---------------------------------------------------------------------------
# tiled direct N-D correlate (corrnd_tiled.cu)
# ---------------------------------------------------------------------------
ND_CASES = [
    # shape, weights shape, origin
    ((40, 64), (3, 3), 0), ((33, 72), (5, 4), (1, -2)), ((7, 16), (9, 9), (0, 3)), ((130, 260), (2, 7), (-1, 0)),
    ((20, 24, 32), (3, 3, 3), 0), ((9, 30, 64), (4, 5, 6), (1, -2, 2)), ((25, 17, 40), (9, 9, 9), 0),
    ((6, 5, 8), (3, 5, 7), (0, 1, -3)), ((10, 12, 64), (2, 3, 21), (0, 1, 4)), ((9, 72), (3, 30), (1, -7)), ((12, 40, 132), (1, 1, 13), (0, 0, 5)), ((18, 18, 20), (7, 1, 2), (3, 0, -1)),
]


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("shape, wshape, origin", ND_CASES)
def test_tiled_correlate_nd(b200, dtype, mode, shape, wshape, origin):
    from dask_image_b200 import _lib

    rng = np.random.default_rng(hash((shape, wshape, mode)) % (2 ** 32))
    a = (rng.random(shape) - 0.25).astype(dtype)
    w = rng.random(wshape) - 0.3
    d = dev(b200, a)
    for name in ("correlate", "convolve"):
        with forced(b200, _lib.FLAG_FORCE_FAST):
            got = getattr(b200.ndimage, name)(d, w, mode=mode, cval=0.6, origin=origin)
        assert _lib.last_kernel().startswith("corrnd_tiled_"), _lib.last_kernel()
        ref = getattr(orc, name)(a, w, mode=mode, cval=0.6, origin=origin)
        assert_parity(got.to_numpy(), ref, what=name, data=a)
    # a tap the reference skips (|w| <= DBL_EPSILON): float32 rows of <= 16 taps stay on the tiled kernel
    # (its zero-skipping instantiation); everything else goes to the scipy-order kernel, which skips too
    w.flat[1] = 0.0
    got = b200.ndimage.correlate(d, w, mode=mode, cval=0.6, origin=origin)
    short_rows = wshape[-1] + 3 <= 16
    if dtype == np.float32 and short_rows:
        assert _lib.last_kernel() == "corrnd_tiled_f32", _lib.last_kernel()
    elif dtype == np.float64:
        assert _lib.last_kernel() == "corrnd_generic", _lib.last_kernel()
    assert_parity(got.to_numpy(), orc.correlate(a, w, mode=mode, cval=0.6, origin=origin), data=a)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_sparse_kernel_on_non_finite_data(b200, dtype):
    """scipy's N-D correlate iterates over the NON-ZERO footprint only, so a zero weight never meets the data:
    a cross-shaped Laplacian (zero corners) next to a NaN / Inf voxel stays finite exactly where scipy's does."""
    from dask_image_b200 import _lib

    rng = np.random.default_rng(77)
    a = rng.random((40, 72)).astype(dtype)
    a[10, 20] = np.nan
    a[25, 60] = np.inf
    a[0, 0] = -np.inf
    w2 = np.array([[0, 1, 0], [1, -4, 1], [0, 1, 0]], dtype=np.float64)
    got = b200.ndimage.correlate(dev(b200, a), w2).to_numpy()
    ref = ndi.correlate(a, w2)
    assert np.array_equal(np.isnan(got), np.isnan(ref)) and np.array_equal(np.isinf(got), np.isinf(ref))
    ok = np.isfinite(ref)
    assert_parity(np.where(ok, got, 0).astype(dtype), np.where(ok, ref, 0).astype(dtype), data=a[np.isfinite(a)])
    assert _lib.last_kernel() == ("corrnd_tiled_f32" if dtype == np.float32 else "corrnd_generic")
    # 3-D, convolve, a whole zero plane and a tiny-but-kept weight
    a3 = rng.random((12, 20, 40)).astype(dtype)
    a3[5, 7, 9] = np.inf
    w3 = rng.random((3, 3, 3)) - 0.5
    w3[0] = 0.0
    w3[1, 1, 2] = 1e-17          # below DBL_EPSILON: skipped by the reference
    got = b200.ndimage.convolve(dev(b200, a3), w3, mode="mirror").to_numpy()
    ref = ndi.convolve(a3, w3, mode="mirror")
    assert np.array_equal(np.isfinite(got), np.isfinite(ref))
    ok = np.isfinite(ref)
    assert_parity(np.where(ok, got, 0).astype(dtype), np.where(ok, ref, 0).astype(dtype), data=a3[np.isfinite(a3)])


def test_config5_convolve_9cubed_wrap(b200):
    """BASELINE.json configs[4] at a checkable size: random 9x9x9 float32 weights,
    mode='wrap', float32 volume, through the public wrapper."""
    from dask_image_b200 import _lib

    rng = np.random.default_rng(1234)
    a = rng.random((48, 56, 96), dtype=np.float32)
    w = rng.random((9, 9, 9), dtype=np.float32)
    got = b200.ndfilters.convolve(dev(b200, a), dev(b200, w), mode="wrap")
    assert _lib.last_kernel() == "corrnd_tiled_f32"
    assert_parity(got.to_numpy(), ndi.convolve(a, w, mode="wrap"))
    # chunk-wise through the dispatch boundary (dask supplies periodic halos)
    chunked = b200.ndfilters.convolve(b200.from_array(a, chunks=(24, 28, 32)), dev(b200, w), mode="wrap")
    assert_parity(np.asarray(chunked), ndi.convolve(a, w, mode="wrap"))


# ---------------------------------------------------------------------------
# tiled minimum / maximum filters (float32, uint8, uint16): bit-exact
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [np.float32, np.uint16, np.uint8])
@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("shape, axis", [((70, 64), 0), ((70, 64), 1), ((5, 130, 48), 1), ((3, 40, 144), 2),
                                         ((260, 512), 0), ((9, 272), 1), ((2, 3, 24, 32), 2), ((300, 16), 0)])
def test_tma_min_max_filter1d(b200, dtype, mode, shape, axis):
    from dask_image_b200 import _lib

    rng = np.random.default_rng(hash((shape, axis, mode)) % (2 ** 32))
    if np.dtype(dtype).kind == "f":
        a = (rng.random(shape) - 0.5).astype(dtype)
        cval = 0.25
    else:
        a = rng.integers(0, np.iinfo(dtype).max, shape, endpoint=True).astype(dtype)
        cval = float(np.iinfo(dtype).max // 2)
    d = dev(b200, a)
    for size, origin in [(2, 0), (3, 1), (7, 0), (7, -3), (8, 3), (16, -8), (31, 4), (40, 0)]:
        for name in ("minimum_filter1d", "maximum_filter1d"):
            # uint8 rows align to 16 elements: up to 15 padding taps, so long windows
            # along the contiguous axis exceed the 48 instantiated taps and fall back
            lead = (-(size // 2 + origin)) % (16 // a.dtype.itemsize)
            tiled = not (axis == a.ndim - 1 and lead + size > 48)
            if tiled:
                with forced(b200, _lib.FLAG_FORCE_FAST):
                    got = getattr(b200.ndimage, name)(d, size, axis=axis, mode=mode, cval=cval, origin=origin)
                assert _lib.last_kernel().startswith(name[:3] + "1d_tma_"), _lib.last_kernel()
            else:
                got = getattr(b200.ndimage, name)(d, size, axis=axis, mode=mode, cval=cval, origin=origin)
                assert _lib.last_kernel() == name[:3] + "1d_generic", _lib.last_kernel()
            ref = getattr(ndi, name)(a, size, axis=axis, mode=mode, cval=cval, origin=origin)
            assert_parity(got.to_numpy(), ref, exact=True, what="%s size=%d origin=%d" % (name, size, origin))


def test_min_max_volume_uses_tiled_kernels(b200):
    from dask_image_b200 import _lib

    a = np.random.default_rng(77).integers(0, 65536, (72, 96, 128), dtype=np.uint16)
    for name in ("minimum_filter", "maximum_filter"):
        _lib.launch_count(reset=True)
        got = getattr(b200.ndfilters, name)(dev(b200, a), size=(5, 3, 7), origin=(1, 0, -2))
        assert _lib.launch_count() == 3 and _lib.last_kernel() == name[:3] + "1d_tma_contig"
        assert_parity(np.asarray(got), getattr(ndi, name)(a, size=(5, 3, 7), origin=(1, 0, -2)), exact=True)
    f = a.astype(np.float32) * 0.37
    got = b200.ndfilters.maximum_filter(dev(b200, f), size=9, mode="mirror")
    assert _lib.last_kernel() == "max1d_tma_contig"
    assert_parity(np.asarray(got), ndi.maximum_filter(f, size=9, mode="mirror"), exact=True)
    # cval that is not a uint16: the exact generic kernel takes the call
    got = b200.ndimage.minimum_filter1d(dev(b200, a), 5, axis=2, mode="constant", cval=-3.5)
    assert _lib.last_kernel() == "min1d_generic"
    assert_parity(got.to_numpy(), ndi.minimum_filter1d(a, 5, axis=2, mode="constant", cval=-3.5), exact=True)
