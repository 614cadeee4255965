"""CPU execution of the exact tiled integer correlate1d kernels against the oracle.

``dask_image_b200/csrc/corr1d_int.cuh`` keeps every phase of its two kernels as a
``__host__ __device__`` function of (thread, block); ``tests/native/int_emul.cu``
runs those functions on the host, block by block.  The arithmetic (the
2^52-biased FMA product, the rotating register windows), the tile geometry,
the boundary staging, the pads of the slab contract and the fused epilogues are
therefore checked bit for bit on the CPU, for every kernel instantiation; the
GPU suite repeats the comparison on the device.
"""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest

from dask_image_b200 import _lib
from oracle import oracle

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "native", "int_emul.cu")
OUT = os.path.join(HERE, "native", "_build", "libint_emul.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
DEPS = [SRC] + [os.path.join(HERE, "..", "dask_image_b200", "csrc", f)
                for f in ("corr1d_int.cuh", "corrnd_int.cuh", "common.cuh")]

MODES = ["reflect", "mirror", "nearest", "wrap", "constant"]
DTYPES = ["uint8", "int8", "uint16", "int16"]


@pytest.fixture(scope="module")
def emul():
    if not (os.path.exists(NVCC) or shutil.which("nvcc")):
        pytest.skip("nvcc is not available")
    if not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in DEPS):
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        subprocess.check_call([NVCC, "-std=c++20", "-O1", "--expt-relaxed-constexpr", "-Wno-deprecated-gpu-targets",
                               "-gencode", "arch=compute_100a,code=sm_100a",
                               "-Xcompiler", "-fPIC,-fno-strict-aliasing", "-shared", "-x", "cu", SRC, "-o", OUT])
    lib = ctypes.CDLL(OUT)
    lib.emul_correlate1d_int.restype = ctypes.c_int
    lib.emul_correlate_nd_int.restype = ctypes.c_int
    return lib


def run_emul(lib, x, w, axis, mode, cval=0.0, origin=0, pad=(0, 0), epilogue=_lib.EPI_STORE, out=None):
    """x: the resident planes (pads included); returns (out over the n planes, info) or (None, None)."""
    x = np.ascontiguousarray(x)
    shape = list(x.shape)
    shape[0] -= pad[0] + pad[1]
    if out is None:
        out = np.zeros(shape, dtype=x.dtype)
    assert out.flags.c_contiguous and out.dtype == x.dtype and list(out.shape) == shape
    w = np.ascontiguousarray(w, dtype=np.float64)
    plane = int(np.prod(shape[1:])) * x.dtype.itemsize
    info = (ctypes.c_int64 * 8)()
    rc = lib.emul_correlate1d_int(
        ctypes.c_void_p(x.ctypes.data + pad[0] * plane), ctypes.c_void_p(out.ctypes.data),
        ctypes.c_int(_lib.DTYPE_CODES[x.dtype.name]), ctypes.c_int(x.ndim), _lib.i64_array(shape),
        ctypes.c_int(axis), w.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), ctypes.c_int64(w.size),
        ctypes.c_int(_lib.mode_code(mode)), ctypes.c_double(cval), ctypes.c_int64(origin),
        ctypes.c_int64(pad[0]), ctypes.c_int64(pad[1]), ctypes.c_int(epilogue), info)
    if rc != 0:
        return None, None
    return out, dict(contig=info[0], kind=info[1], blocks=info[2], smem=info[3], vec=info[4],
                     fixed=info[5], S=info[6], delta=info[7])


def data(rng, shape, dtype):
    ii = np.iinfo(dtype)
    x = rng.integers(ii.min, ii.max + 1, size=shape, dtype=np.int64).astype(dtype)
    # saturated runs exercise the largest pair sums / differences
    flat = x.reshape(-1)
    flat[: flat.size // 7] = ii.max
    flat[flat.size // 7: flat.size // 5] = ii.min
    return x


def weights(kind, nw, rng):
    if kind == "gauss":
        return oracle.gaussian_kernel1d(nw / 8.0 + 0.3, 0, nw // 2)[::-1]
    if kind == "dgauss":
        return oracle.gaussian_kernel1d(nw / 8.0 + 0.3, 1, nw // 2)[::-1]
    if kind == "d2gauss":
        return oracle.gaussian_kernel1d(nw / 8.0 + 0.3, 2, nw // 2)[::-1]
    return rng.standard_normal(nw) * 3.0


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("kind,nw", [("gauss", 17), ("gauss", 3), ("gauss", 1), ("gauss", 11), ("dgauss", 9),
                                     ("dgauss", 15), ("d2gauss", 13), ("random", 8), ("random", 5), ("random", 2),
                                     ("random", 21)])
def test_emulated_kernels_match_oracle(emul, dtype, kind, nw):
    rng = np.random.default_rng(hash((dtype, kind, nw)) % (1 << 32))
    w = weights(kind, nw, rng)
    # aligned and unaligned rows, more than one tile along every axis, ragged ends
    shapes = [(37, 144), (5, 40, 48), (70, 19), (3, 3, 130)]
    for shape in shapes:
        x = data(rng, shape, dtype)
        for axis in range(len(shape)):
            for mode in MODES:
                origins = {0, -(nw // 2), (nw - 1) // 2}
                for origin in sorted(origins):
                    cval = 0.0 if mode != "constant" else float(rng.integers(0, 100))
                    got, info = run_emul(emul, x, w, axis, mode, cval, origin)
                    assert got is not None, "tiled kernel declined an eligible call"
                    assert info["contig"] == int(axis == len(shape) - 1)
                    ref = oracle.correlate1d(x, w, axis=axis, mode=mode, cval=cval, origin=origin)
                    np.testing.assert_array_equal(got, ref, err_msg=str((shape, axis, mode, origin, info)))


@pytest.fixture
def eight_wide(monkeypatch):
    """Run the emulation with eight outputs per lane and walk (the generic NQ form of the walk in
    corr1d_int.cuh; measured slower than four on the B200 -- profiles/r2_int_ab.log -- so the library only
    instantiates four; tests/native/int_emul.cu reads the variable on every call)."""
    monkeypatch.setenv("B200_INT_NQ", "8")


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("kind,nw", [("gauss", 17), ("gauss", 3), ("gauss", 1), ("gauss", 27), ("dgauss", 9),
                                     ("dgauss", 15), ("random", 8), ("random", 5), ("random", 2), ("random", 21)])
def test_emulated_eight_wide_walk_matches_oracle(emul, eight_wide, dtype, kind, nw):
    rng = np.random.default_rng(hash((dtype, kind, nw, 8)) % (1 << 32))
    w = weights(kind, nw, rng)
    for shape in [(37, 144), (5, 40, 48), (70, 19)]:
        x = data(rng, shape, dtype)
        for axis in range(len(shape)):
            for mode in ("reflect", "constant", "wrap"):
                for origin in sorted({0, -(nw // 2), (nw - 1) // 2}):
                    got, info = run_emul(emul, x, w, axis, mode, 5.0, origin)
                    assert got is not None
                    ref = oracle.correlate1d(x, w, axis=axis, mode=mode, cval=5.0, origin=origin)
                    np.testing.assert_array_equal(got, ref, err_msg=str((shape, axis, mode, origin, info)))
    # fused epilogue through the eight-wide store path
    x = data(rng, (24, 48), dtype)
    prev = data(rng, (24, 48), dtype)
    for axis in (0, 1):
        r = oracle.correlate1d(x, w, axis=axis).astype(np.int64)
        out = prev.copy()
        got, _ = run_emul(emul, x, w, axis, "reflect", epilogue=_lib.EPI_SQ_ACC, out=out)
        np.testing.assert_array_equal(got, _wrap(prev.astype(np.int64) + r * r, dtype))


@pytest.mark.parametrize("dtype", ["uint8", "uint16", "int16"])
def test_emulated_kernels_long_filters_and_small_extents(emul, dtype):
    rng = np.random.default_rng(11)
    for nw in (33, 65, 129):
        w = weights("gauss", nw, rng)
        for shape in [(9, 160), (140, 24), (4, 4)]:   # extents far below the filter's reach: multiple reflections
            x = data(rng, shape, dtype)
            for axis in (0, 1):
                for mode in MODES:
                    got, info = run_emul(emul, x, w, axis, mode, 3.0)
                    assert got is not None
                    np.testing.assert_array_equal(got, oracle.correlate1d(x, w, axis=axis, mode=mode, cval=3.0),
                                                  err_msg=str((nw, shape, axis, mode)))


@pytest.mark.parametrize("dtype", DTYPES)
def test_emulated_slab_pads_equal_whole_volume(emul, dtype):
    """pad_lo / pad_hi resident planes: every slab call reproduces the whole-volume call."""
    rng = np.random.default_rng(5)
    x = data(rng, (41, 6, 32), dtype)
    for nw, kind in [(9, "gauss"), (7, "dgauss"), (6, "random")]:
        w = weights(kind, nw, rng)
        for mode in MODES:
            for origin in (0, -1):
                ref = oracle.correlate1d(x, w, axis=0, mode=mode, cval=7.0, origin=origin)
                lo, hi = nw // 2 + origin, nw - 1 - nw // 2 - origin
                bounds = [0, 13, 29, 41]
                for a, b in zip(bounds[:-1], bounds[1:]):
                    pl = 0 if a == 0 else lo
                    ph = 0 if b == 41 else hi
                    if mode == "wrap" and (a == 0 or b == 41):
                        continue   # the slab engine hands wrap slabs their ring neighbours (always padded)
                    got, _ = run_emul(emul, x[a - pl:b + ph], w, 0, mode, 7.0, origin, pad=(pl, ph))
                    assert got is not None
                    np.testing.assert_array_equal(got, ref[a:b], err_msg=str((nw, kind, mode, origin, a, b)))


def _wrap(a, dtype):
    return a.astype(np.int64).astype(dtype)   # modulo 2^bits


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("axis", [0, 1])
def test_emulated_epilogues(emul, dtype, axis):
    """The fused store epilogues of the gradient-magnitude / laplace chains, in the array dtype."""
    rng = np.random.default_rng(3)
    x = data(rng, (24, 48), dtype)
    w = weights("dgauss", 9, rng)
    r = oracle.correlate1d(x, w, axis=axis).astype(np.int64)
    prev = data(rng, (24, 48), dtype)
    p64 = prev.astype(np.int64)
    with np.errstate(invalid="ignore"):
        expect = {
            _lib.EPI_ACC: _wrap(p64 + r, dtype),
            _lib.EPI_SQ_STORE: _wrap(r * r, dtype),
            _lib.EPI_SQ_ACC: _wrap(p64 + r * r, dtype),
            _lib.EPI_SQ_ACC_SQRT: np.sqrt(_wrap(p64 + r * r, dtype)).astype(dtype),
        }
    for epi, ref in expect.items():
        out = prev.copy()
        got, _ = run_emul(emul, x, w, axis, "reflect", epilogue=epi, out=out)
        assert got is not None
        if epi == _lib.EPI_SQ_ACC_SQRT and np.dtype(dtype).kind == "i":
            ok = _wrap(p64 + r * r, dtype) >= 0     # sqrt of a negative: NaN -> cast is platform noise
            np.testing.assert_array_equal(got[ok], ref[ok])
        else:
            np.testing.assert_array_equal(got, ref, err_msg=str(epi))


def test_emulated_kernels_decline_what_they_cannot_hold(emul):
    x = np.zeros((8, 32), dtype=np.uint16)
    w = np.ones(5) / 5
    assert run_emul(emul, x, w, 0, "constant", cval=2.5)[0] is None        # cval is not an array element
    assert run_emul(emul, x, w, 0, "constant", cval=70000.0)[0] is None
    assert run_emul(emul, x, w, 0, "constant", cval=float("nan"))[0] is None
    assert run_emul(emul, x, w, 0, "reflect", cval=2.5)[0] is not None     # cval unused
    assert run_emul(emul, x, np.ones(131), 0, "reflect")[0] is None        # more taps than the parameter block holds
    assert run_emul(emul, x, np.array([1.0, np.inf, 1.0]), 0, "reflect")[0] is None
    assert run_emul(emul, x.astype(np.int32), w, 0, "reflect")[0] is None  # 32-bit sums leave the low word


# ---------------------------------------------------------------------------
# N-D correlate (corrnd_int.cuh)
# ---------------------------------------------------------------------------
def run_emul_nd(lib, x, w, mode, cval=0.0, origin=0, pad=(0, 0)):
    x = np.ascontiguousarray(x)
    shape = list(x.shape)
    shape[0] -= pad[0] + pad[1]
    out = np.zeros(shape, dtype=x.dtype)
    w = np.ascontiguousarray(w, dtype=np.float64)
    origins = [origin] * x.ndim if np.isscalar(origin) else list(origin)
    plane = int(np.prod(shape[1:])) * x.dtype.itemsize
    info = (ctypes.c_int64 * 8)()
    rc = lib.emul_correlate_nd_int(
        ctypes.c_void_p(x.ctypes.data + pad[0] * plane), ctypes.c_void_p(out.ctypes.data),
        ctypes.c_int(_lib.DTYPE_CODES[x.dtype.name]), ctypes.c_int(x.ndim), _lib.i64_array(shape),
        w.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), _lib.i64_array(w.shape), _lib.i64_array(origins),
        ctypes.c_int(_lib.mode_code(mode)), ctypes.c_double(cval), ctypes.c_int64(pad[0]), ctypes.c_int64(pad[1]),
        info)
    if rc != 0:
        return None, None
    return out, dict(blocks=info[0], smem=info[1], vec=info[2], tile_out=info[3])


ND_CASES = [
    ((70, 144), (3, 3)), ((70, 144), (5, 4)), ((37, 50), (1, 7)), ((37, 50), (6, 1)), ((33, 130), (2, 2)),
    ((9, 40, 48), (3, 3, 3)), ((6, 35, 21), (2, 3, 4)), ((5, 34, 160), (1, 5, 5)), ((12, 9, 16), (5, 1, 1)),
    ((3, 4), (5, 7)), ((10, 66, 144), (9, 9, 9)), ((20, 40), (2, 6)),
]


def nd_weights(rng, wshape):
    w = rng.standard_normal(wshape)
    w.reshape(-1)[::5] = 0.0          # taps the reference skips
    w.reshape(-1)[1::7] = 1e-17       # |w| <= DBL_EPSILON: skipped as well
    return w


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape,wshape", ND_CASES)
def test_emulated_nd_kernel_matches_oracle(emul, dtype, shape, wshape):
    rng = np.random.default_rng(hash((dtype, shape, wshape)) % (1 << 32))
    x = data(rng, shape, dtype)
    w = nd_weights(rng, wshape)
    small = np.prod(wshape) < 100
    for mode in (MODES if small else ["reflect", "wrap"]):
        origin_sets = [0, [-(k // 2) for k in wshape], [(k - 1) // 2 for k in wshape]]
        for origin in (origin_sets if small else [0]):
            cval = 0.0 if mode != "constant" else 9.0
            got, info = run_emul_nd(emul, x, w, mode, cval, origin)
            assert got is not None
            ref = oracle.correlate(x, w, mode=mode, cval=cval, origin=origin)
            np.testing.assert_array_equal(got, ref, err_msg=str((mode, origin, info)))


@pytest.mark.parametrize("dtype", ["uint8", "int16"])
@pytest.mark.parametrize("shape,wshape", [((41, 6, 32), (5, 3, 3)), ((41, 48), (4, 3))])
def test_emulated_nd_slab_pads_equal_whole_volume(emul, dtype, shape, wshape):
    rng = np.random.default_rng(6)
    x = data(rng, shape, dtype)
    w = rng.standard_normal(wshape)
    for mode in ["reflect", "mirror", "nearest", "constant"]:
        for origin0 in (0, -1):
            origin = [origin0] + [0] * (len(shape) - 1)
            ref = oracle.correlate(x, w, mode=mode, cval=4.0, origin=origin)
            lo, hi = wshape[0] // 2 + origin0, wshape[0] - 1 - wshape[0] // 2 - origin0
            for a, b in [(0, 13), (13, 29), (29, 41)]:
                pl = 0 if a == 0 else lo
                ph = 0 if b == 41 else hi
                got, _ = run_emul_nd(emul, x[a - pl:b + ph], w, mode, 4.0, origin, pad=(pl, ph))
                assert got is not None
                np.testing.assert_array_equal(got, ref[a:b], err_msg=str((mode, origin0, a, b)))


def test_emulated_nd_kernel_declines(emul):
    x = np.zeros((8, 32), dtype=np.uint16)
    w = np.ones((3, 3))
    assert run_emul_nd(emul, x, w, "constant", cval=0.5)[0] is None
    assert run_emul_nd(emul, x.reshape(2, 4, 8, 4), np.ones((1, 3, 3, 3)), "reflect")[0] is None   # 4-D
    assert run_emul_nd(emul, x.astype(np.uint32), w, "reflect")[0] is None
    assert run_emul_nd(emul, np.zeros((4, 40, 40), np.uint8), np.ones((11, 11, 11)), "reflect")[0] is None


# ---------------------------------------------------------------------------
# randomised: emulated kernels vs scipy itself (what the reference calls per chunk)
# ---------------------------------------------------------------------------
from hypothesis import HealthCheck, given, settings, strategies as st  # noqa: E402
import scipy.ndimage as ndi  # noqa: E402


@settings(max_examples=150, deadline=None, derandomize=True, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(st.data())
def test_emulated_kernels_fuzz_against_scipy(emul, data):
    nd = data.draw(st.integers(1, 3))
    shape = tuple(data.draw(st.sampled_from([1, 2, 3, 5, 8, 16, 31, 40, 67, 130])) for _ in range(nd))
    dt = np.dtype(data.draw(st.sampled_from(DTYPES)))
    mode = data.draw(st.sampled_from(MODES))
    cval = float(data.draw(st.sampled_from([0, 1, 100, -7]))) if mode == "constant" else 0.0
    if dt.kind == "u" and cval < 0:
        cval = 7.0
    rng = np.random.default_rng(data.draw(st.integers(0, 2 ** 16)))
    ii = np.iinfo(dt)
    a = rng.integers(ii.min, ii.max + 1, size=shape, dtype=np.int64).astype(dt)
    axis = data.draw(st.integers(0, nd - 1))
    nw = data.draw(st.integers(1, 40))
    origin = data.draw(st.integers(-(nw // 2), (nw - 1) // 2))
    w = rng.standard_normal(nw)
    if nw % 2 and data.draw(st.booleans()):
        w = (w + w[::-1]) / 2 if data.draw(st.booleans()) else (w - w[::-1]) / 2     # the folded branches
    got, info = run_emul(emul, a, w, axis, mode, cval, origin)
    assert got is not None
    np.testing.assert_array_equal(got, ndi.correlate1d(a, w, axis=axis, mode=mode, cval=cval, origin=origin),
                                  err_msg=str((shape, dt, axis, mode, nw, origin, info)))
    if nd >= 2:
        wshape = tuple(data.draw(st.integers(1, 5)) for _ in range(nd))
        wn = rng.standard_normal(wshape)
        origins = [data.draw(st.integers(-(s // 2), (s - 1) // 2)) for s in wshape]
        got, info = run_emul_nd(emul, a, wn, mode, cval, origins)
        assert got is not None
        np.testing.assert_array_equal(got, ndi.correlate(a, wn, mode=mode, cval=cval, origin=origins),
                                      err_msg=str((shape, dt, wshape, mode, origins, info)))
