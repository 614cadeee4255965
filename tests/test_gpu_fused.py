"""GPU tests of the fused pass pair (corr2d_fused.cu): the passes along the last
two axes of a float32 array in one launch.  The contract is stronger than the
tolerance: the fused launch is BIT-IDENTICAL to the two one-pass kernels it
replaces (same operations in the same order; the intermediate is rounded to
float32 in shared memory exactly as the stored pass would be), so the fused and
unfused paths of every public function must agree with ``array_equal`` -- on
top of both matching the oracle within the north_star tolerance (fp32 1e-5)."""
import contextlib

import numpy as np
import pytest

from _helpers import assert_parity
from oracle import oracle as orc

pytestmark = pytest.mark.gpu

MODES = ["reflect", "wrap", "nearest", "constant", "mirror"]


@pytest.fixture(scope="module")
def b200():
    import dask_image_b200

    return dask_image_b200


@contextlib.contextmanager
def fusion(b200, on):
    old = b200.ndimage.set_pair_fusion(on)
    try:
        yield
    finally:
        b200.ndimage.set_pair_fusion(old)


def dev(b200, a):
    return b200.B200Array.from_numpy(np.asarray(a))


def both(b200, fn, *args, **kw):
    """(fused result, kernels it launched, unfused result)"""
    from dask_image_b200 import _lib

    with fusion(b200, True):
        _lib.launch_count(reset=True)
        fused = fn(*args, **kw).to_numpy()
        n_fused = _lib.launch_count()
        k = _lib.last_kernel()
    with fusion(b200, False):
        _lib.launch_count(reset=True)
        plain = fn(*args, **kw).to_numpy()
        n_plain = _lib.launch_count()
    return fused, k, n_fused, plain, n_plain


# tile-edge heavy shapes: fewer rows / columns than a tile, ragged last tiles in both directions,
# several planes, 2-D (a single launch) and 4-D (leading axes collapse)
SHAPES = [(40, 64), (33, 260), (70, 516), (3, 37, 100), (2, 130, 244), (5, 8, 8), (2, 2, 65, 300), (1, 300, 12)]
SIGMAS = [0.3, 0.5, 1.0, 1.5, 2.0, 2.5, 3.0, 3.5, 4.0, 5.0]


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("shape", SHAPES)
def test_fused_gaussian_bit_identical_to_two_passes(b200, shape, mode):
    rng = np.random.default_rng(hash((shape, mode)) % (2 ** 32))
    a = (rng.random(shape) * 3 - 1).astype(np.float32)
    d = dev(b200, a)
    for sigma in SIGMAS:
        fused, k, nf, plain, npl = both(b200, b200.ndimage.gaussian_filter, d, sigma, mode=mode, cval=0.75)
        assert k.startswith("corr2d_fused_tma"), (k, sigma)
        assert nf == npl - 1 and nf == len(shape) - 1
        assert np.array_equal(fused, plain), "sigma=%s: %d voxels differ" % (sigma, int((fused != plain).sum()))
        assert_parity(fused, orc.gaussian_filter(a, sigma, mode=mode, cval=0.75), what="sigma=%s" % sigma)


@pytest.mark.parametrize("mode", ["reflect", "constant", "wrap"])
def test_fused_derivative_chains(b200, mode):
    """gradient magnitude / laplace: every chain ends in a fused pair with an accumulating epilogue; orders differ
    per axis, radii do not."""
    from dask_image_b200 import _lib

    rng = np.random.default_rng(5)
    a = rng.random((20, 70, 260)).astype(np.float32)
    d = dev(b200, a)
    for sigma in (1.0, 2.0, 3.0):
        for fn, ref in ((b200.ndimage.gaussian_gradient_magnitude, orc.gaussian_gradient_magnitude),
                        (b200.ndimage.gaussian_laplace, orc.gaussian_laplace)):
            fused, k, nf, plain, npl = both(b200, fn, d, sigma, mode=mode, cval=-0.5)
            assert (nf, npl) == (5, 8), (nf, npl)
            assert np.array_equal(fused, plain)
            assert_parity(fused, ref(a, sigma, mode=mode, cval=-0.5), data=a)
        for order in [(0, 1, 0), (0, 0, 2), (1, 1, 1), (0, 3, 1)]:
            fused, k, nf, plain, npl = both(b200, b200.ndimage.gaussian_filter, d, sigma, order=order, mode=mode)
            assert np.array_equal(fused, plain)
            assert_parity(fused, orc.gaussian_filter(a, sigma, order=order, mode=mode), data=a)
    assert _lib.last_kernel()


def test_fused_per_axis_modes_and_constant_cval(b200):
    """mode 'constant' with cval != 0 along the contiguous axis: the second pass must see cval beyond the ends
    of the STORED first pass (not the first pass of a cval column) -- derivative taps make the two differ."""
    rng = np.random.default_rng(6)
    a = rng.random((70, 300)).astype(np.float32)
    d = dev(b200, a)
    for modes in [("reflect", "constant"), ("constant", "mirror"), ("wrap", "nearest"), ("constant", "constant")]:
        for order in (0, (1, 0), (0, 1), (2, 1)):
            fused, k, nf, plain, npl = both(b200, b200.ndimage.gaussian_filter, d, 2.0, order=order, mode=modes,
                                            cval=3.5)
            assert k.startswith("corr2d_fused") and nf == 1
            assert np.array_equal(fused, plain), (modes, order)
            assert_parity(fused, orc.gaussian_filter(a, 2.0, order=order, mode=list(modes), cval=3.5), data=a)


def test_fused_non_finite_stays_local(b200):
    """A NaN / Inf voxel poisons exactly the window around it, as in the two-pass path."""
    a = np.random.default_rng(7).random((4, 90, 280)).astype(np.float32)
    a[1, 40, 100] = np.nan
    a[2, 5, 270] = np.inf
    fused, k, nf, plain, npl = both(b200, b200.ndimage.gaussian_filter, dev(b200, a), 1.5)
    assert np.array_equal(fused, plain, equal_nan=True)
    ref = orc.gaussian_filter(a, 1.5)
    assert np.array_equal(np.isfinite(fused), np.isfinite(ref))


@pytest.mark.parametrize("nx", [1, 2, 3, 5, 7, 13, 31, 67, 110, 255, 257, 1023, 1034])
def test_fused_unaligned_rows_take_the_ldg_staging(b200, nx):
    """Rows that are not 16-byte multiples cannot be TMA-staged; the same tiles are filled with plain loads
    and the arithmetic is unchanged -- no fall back to the one-thread-per-output kernel."""
    from dask_image_b200 import _lib

    rng = np.random.default_rng(nx)
    a = rng.random((3, 45, nx)).astype(np.float32)
    d = dev(b200, a)
    for sigma, mode in ((1.0, "reflect"), (2.0, "mirror"), (3.0, "wrap"), (2.0, "constant")):
        with fusion(b200, True):
            got = b200.ndimage.gaussian_filter(d, sigma, mode=mode, cval=1.25).to_numpy()
        k = _lib.last_kernel()
        assert k == ("corr2d_fused_tma_f32" if nx % 4 == 0 else "corr2d_fused_ldg_f32"), (nx, k)
        assert_parity(got, orc.gaussian_filter(a, sigma, mode=mode, cval=1.25), what="nx=%d sigma=%s" % (nx, sigma))


def test_pair_abi_direct_and_not_eligible(b200):
    """The C entry point itself: bit-identical to two b200_correlate1d calls; filters without a fused kernel
    are refused with 'not eligible' (the caller then launches the passes one by one)."""
    import torch

    from dask_image_b200 import _lib, ndimage

    lib = _lib.load()
    rng = np.random.default_rng(8)
    a = rng.random((6, 50, 272)).astype(np.float32)
    x = torch.from_numpy(a).cuda()
    w = ndimage._gaussian_weights(2.0, 0, 4.0)
    ps1, ps2 = ndimage._Pass("corr", 1, "reflect", weights=w), ndimage._Pass("corr", 2, "mirror", weights=w)
    out = torch.empty_like(x)
    ndimage._launch_correlate1d_pair(x, out, a.shape, a.dtype, ps1, ps2, 0.0)
    tmp, ref = torch.empty_like(x), torch.empty_like(x)
    ndimage._launch_correlate1d(x, tmp, a.shape, a.dtype, 1, w, "reflect", 0.0, 0)
    ndimage._launch_correlate1d(tmp, ref, a.shape, a.dtype, 2, w, "mirror", 0.0, 0)
    assert torch.equal(out, ref)
    assert lib.b200_correlate1d_pair_eligible(8, 3, _lib.i64_array(a.shape), 17, 0, 19, 0) == 0
    assert lib.b200_correlate1d_pair_eligible(9, 3, _lib.i64_array(a.shape), 17, 0, 17, 0) == 0   # float64
    assert lib.b200_correlate1d_pair_eligible(8, 1, _lib.i64_array(a.shape[:1]), 17, 0, 17, 0) == 0
    bad = ndimage._Pass("corr", 2, "mirror", weights=np.ones(19))
    with pytest.raises(RuntimeError, match="not eligible"):
        ndimage._launch_correlate1d_pair(x, out, a.shape, a.dtype, ps1, bad, 0.0)


# ---- the exact integer box filter along the last two axes in one launch -----------------------------
# rows are 16-byte multiples (the pair is TMA-staged only); ragged tiles in both directions
BOX_SHAPES = {
    np.uint16: [(40, 64), (33, 264), (70, 520), (3, 37, 104), (2, 130, 248), (5, 8, 8), (2, 2, 65, 304), (1, 300, 16)],
    np.uint8: [(40, 64), (33, 272), (70, 528), (3, 37, 112), (2, 130, 256), (5, 16, 16), (2, 2, 65, 304)],
}


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("dtype", [np.uint16, np.uint8])
def test_fused_box_pair_bit_exact(b200, dtype, mode):
    """uniform_filter on u8 / u16: the box passes along the last two axes run as ONE launch and the result is
    bit-exact against scipy's (oracle) and equal to the pass-by-pass path."""
    import scipy.ndimage as ndi

    info = np.iinfo(dtype)
    for shape in BOX_SHAPES[dtype]:
        rng = np.random.default_rng(hash((shape, mode, info.max)) % (2 ** 32))
        a = rng.integers(0, info.max + 1, shape).astype(dtype)
        a.flat[:: 7] = info.max          # the largest window sums
        d = dev(b200, a)
        for size in (3, 5, 7, 9, 11, 13, 15, 17):
            cval = 7.0 if mode == "constant" else 0.0
            fused, k, nf, plain, npl = both(b200, b200.ndimage.uniform_filter, d, size, mode=mode, cval=cval)
            assert k == "uniform2d_fused_tma_" + np.dtype(dtype).name.replace("int", ""), (k, shape, size)
            assert nf == npl - 1 and nf == len(shape) - 1
            ref = ndi.uniform_filter(a, size, mode=mode, cval=cval)
            assert np.array_equal(fused, plain), (shape, size, int((fused != plain).sum()))
            assert_parity(fused, ref, exact=True, what="%s size=%d" % (shape, size))


def test_fused_box_pair_direct_form_equals_running_sums(b200, monkeypatch):
    rng = np.random.default_rng(11)
    a = rng.integers(0, 65536, (3, 70, 520)).astype(np.uint16)
    d = dev(b200, a)
    want = b200.ndimage.uniform_filter(d, 7).to_numpy()
    monkeypatch.setenv("B200_BOX2D_SLIDE", "0")
    got = b200.ndimage.uniform_filter(d, 7).to_numpy()
    assert np.array_equal(got, want)


def test_fused_box_pair_eligibility(b200):
    """What does NOT take the pair (and still gives scipy's values pass by pass): even sizes, origins, unequal
    sizes, rows that are not 16-byte multiples, floats, signed types, constant mode with a non-integer cval."""
    import scipy.ndimage as ndi

    from dask_image_b200 import _lib

    rng = np.random.default_rng(5)
    a = rng.integers(0, 65536, (3, 40, 64)).astype(np.uint16)
    cases = [
        (a, dict(size=4)), (a, dict(size=7, origin=1)), (a, dict(size=(3, 7, 5))),
        (a[:, :, :63].copy(), dict(size=7)), (a.astype(np.int16), dict(size=7)),
        (a, dict(size=7, mode="constant", cval=0.5)),
    ]
    for arr, kw in cases:
        got = b200.ndimage.uniform_filter(dev(b200, arr), **kw).to_numpy()
        assert not _lib.last_kernel().startswith("uniform2d_fused"), kw
        assert_parity(got, ndi.uniform_filter(arr, **kw), exact=True, what=str(kw))
    # sizes (1, 7, 7): the first axis is skipped, the pair still applies; (7, 7, 1) leaves no pair
    got = b200.ndimage.uniform_filter(dev(b200, a), (1, 7, 7)).to_numpy()
    assert _lib.last_kernel() == "uniform2d_fused_tma_u16"
    assert_parity(got, ndi.uniform_filter(a, (1, 7, 7)), exact=True)
    lib = _lib.load()
    sh = _lib.i64_array(a.shape)
    assert lib.b200_uniform_filter1d_pair_eligible(2, 3, sh, 7, 0, 2, 7, 0, 2, 0.0) == 1
    assert lib.b200_uniform_filter1d_pair_eligible(8, 3, sh, 7, 0, 2, 7, 0, 2, 0.0) == 0      # float32
    assert lib.b200_uniform_filter1d_pair_eligible(2, 3, sh, 7, 0, 2, 9, 0, 2, 0.0) == 0
    assert lib.b200_uniform_filter1d_pair_eligible(2, 3, sh, 19, 0, 2, 19, 0, 2, 0.0) == 0    # not compiled
    assert lib.b200_uniform_filter1d_pair_eligible(2, 1, _lib.i64_array(a.shape[:1]), 7, 0, 2, 7, 0, 2, 0.0) == 0


def test_fused_box_pair_unaligned_base_pointer_takes_two_launches(b200):
    import scipy.ndimage as ndi
    import torch

    from dask_image_b200 import _lib

    rng = np.random.default_rng(6)
    a = rng.integers(0, 65536, (2, 40, 64)).astype(np.uint16)
    flat = torch.from_numpy(np.concatenate([np.zeros(1, np.uint16), a.reshape(-1)])).cuda()
    x = b200.B200Array(flat[1:].view(a.shape))             # base pointer 2 bytes off a 16-byte boundary
    assert x.tensor.data_ptr() % 16 == 2
    got = b200.ndimage.uniform_filter(x, 7).to_numpy()
    assert not _lib.last_kernel().startswith("uniform2d_fused")
    assert_parity(got, ndi.uniform_filter(a, 7), exact=True)


def test_fused_box_pair_in_sharded_and_host_volumes(b200):
    """The same pair inside the slab engine (one rank) and the streamed host volume."""
    import scipy.ndimage as ndi

    rng = np.random.default_rng(8)
    a = rng.integers(0, 65536, (24, 72, 136)).astype(np.uint16)
    ref = ndi.uniform_filter(a, 7)
    v = b200.ShardedVolume.from_global(a)
    assert_parity(b200.ndfilters.uniform_filter(v, 7).gather(), ref, exact=True)
    h = b200.HostVolume(a, chunk_bytes=8 * 72 * 136 * 2)
    assert_parity(b200.ndfilters.uniform_filter(h, 7).to_numpy(), ref, exact=True)


@pytest.mark.parametrize("mode", ["reflect", "constant", "wrap"])
def test_fused_float_box_pair(b200, mode):
    """float32 uniform_filter: the box filter is a correlation with taps 1 / size, so its last two passes take the
    float32 pair kernel -- bit-identical to the pass-by-pass path, within the fp32 tolerance of scipy."""
    import scipy.ndimage as ndi

    for shape in [(40, 64), (3, 37, 100), (2, 130, 244), (2, 33, 63)]:
        rng = np.random.default_rng(hash((shape, mode)) % (2 ** 32))
        a = (rng.random(shape) * 5 - 1).astype(np.float32)
        d = dev(b200, a)
        for size in (3, 5, 7, 9, 15, 21, 33):
            fused, k, nf, plain, npl = both(b200, b200.ndimage.uniform_filter, d, size, mode=mode, cval=0.3)
            assert k.startswith("corr2d_fused_"), (k, shape, size)
            assert nf == npl - 1 and nf == len(shape) - 1
            assert np.array_equal(fused, plain), (shape, size, int((fused != plain).sum()))
            assert_parity(fused, ndi.uniform_filter(a, size, mode=mode, cval=0.3), what="%s size=%d" % (shape, size),
                          data=a)
    # even sizes / origins / unequal sizes: no pair
    from dask_image_b200 import _lib

    a3 = np.random.default_rng(4).random((4, 40, 64)).astype(np.float32)
    for kw in (dict(size=4), dict(size=7, origin=1), dict(size=(3, 7, 5))):
        got = b200.ndimage.uniform_filter(dev(b200, a3), **kw).to_numpy()
        assert not _lib.last_kernel().startswith("corr2d_fused"), kw
        assert_parity(got, ndi.uniform_filter(a3, **kw), data=a3)
