"""GPU tests of the exact tiled integer correlate1d kernels (corr1d_int.cu):
bit-for-bit against the CPU oracle, for every instantiation (4 dtypes x
symmetric / antisymmetric / general taps x strided / contiguous axis), ragged and
unaligned shapes, every boundary mode, origins, the slab pads and the fused
epilogues -- the same matrix tests/test_int_emul.py runs through the host
emulation of these kernels."""
import numpy as np
import pytest

from oracle import oracle as orc
from test_int_emul import DTYPES, MODES, data, weights, _wrap

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def b200():
    import dask_image_b200

    return dask_image_b200


def dev(b200, a):
    return b200.B200Array.from_numpy(np.asarray(a))


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("kind,nw", [("gauss", 17), ("gauss", 3), ("gauss", 1), ("gauss", 11), ("dgauss", 9),
                                     ("dgauss", 15), ("d2gauss", 13), ("random", 8), ("random", 5), ("random", 2),
                                     ("random", 21), ("gauss", 65), ("gauss", 129)])
def test_int_correlate1d_bitwise(b200, dtype, kind, nw):
    from dask_image_b200 import _lib

    rng = np.random.default_rng(hash((dtype, kind, nw)) % (1 << 32))
    w = weights(kind, nw, rng)
    for shape in [(37, 144), (5, 40, 48), (70, 19), (3, 3, 130), (300, 272), (4, 4)]:
        x = data(rng, shape, dtype)
        d = dev(b200, x)
        for axis in range(len(shape)):
            for mode in MODES:
                for origin in sorted({0, -(nw // 2), (nw - 1) // 2}):
                    cval = 0.0 if mode != "constant" else float(rng.integers(0, 100))
                    got = b200.ndimage.correlate1d(d, w, axis=axis, mode=mode, cval=cval, origin=origin)
                    want = "corr1d_int_contig" if axis == len(shape) - 1 else "corr1d_int_strided"
                    assert _lib.last_kernel() == want, _lib.last_kernel()
                    ref = orc.correlate1d(x, w, axis=axis, mode=mode, cval=cval, origin=origin)
                    np.testing.assert_array_equal(got.to_numpy(), ref, err_msg=str((shape, axis, mode, origin)))


@pytest.mark.parametrize("dtype", DTYPES)
def test_int_tiled_equals_generic_kernel(b200, dtype):
    """Same bits as the one-thread-per-output kernel the tiled path replaces (B200_FLAG_EXACT)."""
    from dask_image_b200 import _lib

    rng = np.random.default_rng(8)
    x = data(rng, (96, 200, 64), dtype)
    d = dev(b200, x)
    for sigma, order in [(1.0, 0), (2.0, 0), (2.0, 1), (3.0, 2)]:
        for axis in (0, 1, 2):
            fast = b200.ndimage.gaussian_filter1d(d, sigma, axis=axis, order=order).to_numpy()
            assert _lib.last_kernel().startswith("corr1d_int_")
            old = b200.ndimage.set_default_flags(_lib.FLAG_EXACT)
            try:
                exact = b200.ndimage.gaussian_filter1d(d, sigma, axis=axis, order=order).to_numpy()
                assert _lib.last_kernel() == "corr1d_generic"
            finally:
                b200.ndimage.set_default_flags(old)
            np.testing.assert_array_equal(fast, exact)


@pytest.mark.parametrize("dtype", DTYPES)
def test_int_slab_pads_equal_whole_volume(b200, dtype):
    import torch
    from dask_image_b200 import _lib, ndimage

    rng = np.random.default_rng(5)
    x = data(rng, (41, 6, 32), dtype)
    xt = torch.from_numpy(x.view(np.uint8).reshape(x.shape + (x.dtype.itemsize,))).cuda()
    plane = 6 * 32 * x.dtype.itemsize
    for nw, kind in [(9, "gauss"), (7, "dgauss"), (6, "random")]:
        w = weights(kind, nw, rng)
        for mode in ["reflect", "mirror", "nearest", "constant"]:
            for origin in (0, -1):
                ref = orc.correlate1d(x, w, axis=0, mode=mode, cval=7.0, origin=origin)
                lo, hi = nw // 2 + origin, nw - 1 - nw // 2 - origin
                for a, b in [(0, 13), (13, 29), (29, 41)]:
                    pl = 0 if a == 0 else lo
                    ph = 0 if b == 41 else hi
                    src = xt.reshape(-1)[a * plane:]
                    out = torch.empty((b - a) * plane, dtype=torch.uint8, device="cuda")
                    ndimage._launch_correlate1d(src, out, (b - a, 6, 32), x.dtype, 0, w, mode, 7.0, origin,
                                                pad=(pl, ph))
                    assert _lib.last_kernel() == "corr1d_int_strided"
                    got = out.cpu().numpy().view(x.dtype).reshape(b - a, 6, 32)
                    np.testing.assert_array_equal(got, ref[a:b], err_msg=str((nw, kind, mode, origin, a, b)))


@pytest.mark.parametrize("dtype", DTYPES)
def test_int_chains_bitwise(b200, dtype):
    """gaussian_filter / gradient magnitude / laplace / sobel on integer images: every pass on the tiled
    kernels (fused epilogues included), bit-identical to the oracle."""
    from dask_image_b200 import _lib

    rng = np.random.default_rng(21)
    for shape in [(40, 48, 64), (130, 144)]:
        x = data(rng, shape, dtype) // 4
        d = dev(b200, x)
        for mode in ("reflect", "constant", "wrap"):
            np.testing.assert_array_equal(b200.ndimage.gaussian_filter(d, 1.5, mode=mode).to_numpy(),
                                          orc.gaussian_filter(x, 1.5, mode=mode))
            assert _lib.last_kernel().startswith("corr1d_int_")
            np.testing.assert_array_equal(b200.ndimage.gaussian_gradient_magnitude(d, 1.0, mode=mode).to_numpy(),
                                          orc.gaussian_gradient_magnitude(x, 1.0, mode=mode))
            np.testing.assert_array_equal(b200.ndimage.gaussian_laplace(d, 1.0, mode=mode).to_numpy(),
                                          orc.gaussian_laplace(x, 1.0, mode=mode))


# ---------------------------------------------------------------------------
# N-D correlate / convolve (corrnd_int.cu)
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", DTYPES)
def test_int_correlate_nd_bitwise(b200, dtype):
    from dask_image_b200 import _lib
    from test_int_emul import ND_CASES, nd_weights

    for shape, wshape in ND_CASES + [((300, 272), (3, 3)), ((20, 70, 256), (3, 3, 3))]:
        rng = np.random.default_rng(hash((dtype, shape, wshape)) % (1 << 32))
        x = data(rng, shape, dtype)
        d = dev(b200, x)
        w = nd_weights(rng, wshape)
        small = np.prod(wshape) < 100
        for mode in (MODES if small else ["reflect", "wrap"]):
            origin_sets = [0, [-(k // 2) for k in wshape], [(k - 1) // 2 for k in wshape]]
            for origin in (origin_sets if small else [0]):
                cval = 0.0 if mode != "constant" else 9.0
                got = b200.ndimage.correlate(d, w, mode=mode, cval=cval, origin=origin)
                assert _lib.last_kernel() == "corrnd_int", _lib.last_kernel()
                ref = orc.correlate(x, w, mode=mode, cval=cval, origin=origin)
                np.testing.assert_array_equal(got.to_numpy(), ref, err_msg=str((shape, wshape, mode, origin)))
        got = b200.ndimage.convolve(d, w, mode="mirror")
        np.testing.assert_array_equal(got.to_numpy(), orc.convolve(x, w, mode="mirror"))


@pytest.mark.parametrize("dtype", ["uint8", "int16"])
def test_int_correlate_nd_equals_generic_kernel(b200, dtype):
    from dask_image_b200 import _lib

    rng = np.random.default_rng(12)
    x = data(rng, (48, 96, 128), dtype)
    d = dev(b200, x)
    w = rng.standard_normal((3, 5, 3))
    fast = b200.ndimage.correlate(d, w, mode="nearest").to_numpy()
    assert _lib.last_kernel() == "corrnd_int"
    old = b200.ndimage.set_default_flags(_lib.FLAG_EXACT)
    try:
        exact = b200.ndimage.correlate(d, w, mode="nearest").to_numpy()
        assert _lib.last_kernel() == "corrnd_generic"
    finally:
        b200.ndimage.set_default_flags(old)
    np.testing.assert_array_equal(fast, exact)
