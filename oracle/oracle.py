"""CPU oracle for the dask-image ndfilters hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline leg
may import this module (and ``oracle/reference_arm.py``, the reference's own public API for that leg); the product
package (``dask_image_b200``) never does.

Layers restated here (parity status: PINNED -- against scipy 1.18.1 run here, ``tests/test_oracle.py`` and
``tests/golden/``, and against the REFERENCE ITSELF run here: the unmodified ``dask_image.ndfilters`` imports with a
stand-in for ``dask.utils.Dispatch``, the one name it needs from dask, and reproduces every golden output and this
module's ``map_overlap_emulated`` results bit for bit, ``tests/test_reference_wrappers_cpu.py``):

* the three compiled loops of scipy's ``_nd_image`` module -> ``ni_oracle.c``
  (loaded through ctypes below);
* scipy's Python drivers that turn ``gaussian_filter`` & co. into chains of
  1-D passes (``scipy/ndimage/_filters.py``, a third-party dependency of the
  reference, ``pyproject.toml:29``) -> plain numpy in this file;
* dask's ``map_overlap`` chunk/halo/trim step that dask-image wraps around the
  per-chunk call (``dask_image/ndfilters/_gaussian.py:70-81`` etc.) ->
  ``map_overlap_emulated`` (dask is not installable in this image).
"""
import ctypes
import math
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_DT = {
    np.dtype(np.uint8): 0, np.dtype(np.int8): 1, np.dtype(np.uint16): 2,
    np.dtype(np.int16): 3, np.dtype(np.uint32): 4, np.dtype(np.int32): 5,
    np.dtype(np.uint64): 6, np.dtype(np.int64): 7, np.dtype(np.float32): 8,
    np.dtype(np.float64): 9,
}
_MODES = {"nearest": 0, "wrap": 1, "grid-wrap": 1, "reflect": 2,
          "grid-mirror": 2, "mirror": 3, "constant": 4, "grid-constant": 4}


def build(force=False):
    """Compile ni_oracle.c with the committed Makefile (gcc only)."""
    so = os.path.join(_HERE, "libni_oracle.so")
    src = os.path.join(_HERE, "ni_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return so


def _lib():
    global _LIB
    if _LIB is None:
        lib = ctypes.CDLL(build())
        i64p = ctypes.POINTER(ctypes.c_int64)
        f64p = ctypes.POINTER(ctypes.c_double)
        lib.oracle_correlate1d.argtypes = [
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, i64p,
            ctypes.c_int, f64p, ctypes.c_int64, ctypes.c_int, ctypes.c_double,
            ctypes.c_int64]
        lib.oracle_uniform_filter1d.argtypes = [
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, i64p,
            ctypes.c_int, ctypes.c_int64, ctypes.c_int, ctypes.c_double,
            ctypes.c_int64]
        lib.oracle_correlate_nd.argtypes = [
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, i64p,
            f64p, i64p, i64p, ctypes.c_int, ctypes.c_double]
        _LIB = lib
    return _LIB


def _mode_code(mode):
    try:
        return _MODES[mode]
    except (KeyError, TypeError):
        # scipy/ndimage/_ni_support.py:59
        raise RuntimeError("boundary mode not supported")


def _prep(a):
    a = np.ascontiguousarray(a)
    if a.dtype not in _DT:
        raise RuntimeError("data type not supported")
    return a


def _i64(seq):
    return (ctypes.c_int64 * len(seq))(*[int(s) for s in seq])


def _check(rc, what):
    if rc == 2:
        raise ValueError("Invalid origin")
    if rc != 0:
        raise RuntimeError("%s failed (%d)" % (what, rc))


# --------------------------------------------------------------------------
# the three native loops
# --------------------------------------------------------------------------
def correlate1d(input, weights, axis=-1, mode="reflect", cval=0.0, origin=0):
    """scipy.ndimage.correlate1d (scipy/ndimage/_filters.py:556-612)."""
    a = _prep(input)
    w = np.ascontiguousarray(weights, dtype=np.float64)
    if w.ndim != 1 or w.shape[0] < 1:
        raise RuntimeError("no filter weights given")
    axis = axis + a.ndim if axis < 0 else axis
    out = np.empty_like(a)
    rc = _lib().oracle_correlate1d(
        a.ctypes.data, out.ctypes.data, _DT[a.dtype], a.ndim, _i64(a.shape), axis,
        w.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), w.shape[0],
        _mode_code(mode), float(cval), int(origin))
    _check(rc, "correlate1d")
    return out


def uniform_filter1d(input, size, axis=-1, mode="reflect", cval=0.0, origin=0):
    """scipy.ndimage.uniform_filter1d (scipy/ndimage/_filters.py:1502-1549)."""
    a = _prep(input)
    if size < 1:
        raise RuntimeError("incorrect filter size")
    axis = axis + a.ndim if axis < 0 else axis
    out = np.empty_like(a)
    rc = _lib().oracle_uniform_filter1d(
        a.ctypes.data, out.ctypes.data, _DT[a.dtype], a.ndim, _i64(a.shape), axis,
        int(size), _mode_code(mode), float(cval), int(origin))
    _check(rc, "uniform_filter1d")
    return out


def _correlate_or_convolve(input, weights, mode, cval, origin, convolution):
    """scipy/ndimage/_filters.py:1253-1309."""
    a = _prep(input)
    w = np.asarray(weights, dtype=np.float64)
    if w.ndim != a.ndim or any(s < 1 for s in w.shape):
        raise RuntimeError("weights.ndim must match the input's")
    origins = [int(origin)] * a.ndim if np.isscalar(origin) else [int(o) for o in origin]
    if convolution:
        w = w[tuple([slice(None, None, -1)] * w.ndim)]
        for i in range(a.ndim):
            origins[i] = -origins[i]
            if not w.shape[i] & 1:
                origins[i] -= 1
    w = np.ascontiguousarray(w)
    out = np.empty_like(a)
    rc = _lib().oracle_correlate_nd(
        a.ctypes.data, out.ctypes.data, _DT[a.dtype], a.ndim, _i64(a.shape),
        w.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), _i64(w.shape),
        _i64(origins), _mode_code(mode), float(cval))
    _check(rc, "correlate")
    return out


def correlate(input, weights, mode="reflect", cval=0.0, origin=0):
    return _correlate_or_convolve(input, weights, mode, cval, origin, False)


def convolve(input, weights, mode="reflect", cval=0.0, origin=0):
    return _correlate_or_convolve(input, weights, mode, cval, origin, True)


# --------------------------------------------------------------------------
# scipy's Python drivers, restated
# --------------------------------------------------------------------------
def gaussian_kernel1d(sigma, order, radius):
    """scipy/ndimage/_filters.py:656-684 -- the operator-matrix construction
    (closed forms agree only to ~1 ulp)."""
    if order < 0:
        raise ValueError("order must be non-negative")
    exponent_range = np.arange(order + 1)
    sigma2 = sigma * sigma
    x = np.arange(-radius, radius + 1)
    phi_x = np.exp(-0.5 / sigma2 * x ** 2)
    phi_x = phi_x / phi_x.sum()
    if order == 0:
        return phi_x
    q = np.zeros(order + 1)
    q[0] = 1
    D = np.diag(exponent_range[1:], 1)
    P = np.diag(np.ones(order) / -sigma2, -1)
    Q_deriv = D + P
    for _ in range(order):
        q = Q_deriv.dot(q)
    q = (x[:, None] ** exponent_range).dot(q)
    return q * phi_x


def gaussian_filter1d(input, sigma, axis=-1, order=0, mode="reflect", cval=0.0,
                      truncate=4.0):
    """scipy/ndimage/_filters.py:688-754; kernel reversed before correlate (:753)."""
    sd = float(sigma)
    lw = int(truncate * sd + 0.5)
    weights = gaussian_kernel1d(sigma, order, lw)[::-1]
    return correlate1d(input, weights, axis, mode, cval, 0)


def _seq(v, n):
    if isinstance(v, str) or not np.iterable(v):
        return [v] * n
    v = list(v)
    if len(v) != n:
        raise RuntimeError("sequence argument must have length equal to input rank")
    return v


def gaussian_filter(input, sigma, order=0, mode="reflect", cval=0.0, truncate=4.0):
    """scipy/ndimage/_filters.py:758-862: one 1-D pass per axis with
    sigma > 1e-15, in axis order, each pass stored in the array dtype."""
    a = _prep(input)
    orders = _seq(order, a.ndim)
    sigmas = _seq(sigma, a.ndim)
    modes = _seq(mode, a.ndim)
    out = a
    done = False
    for ax in range(a.ndim):
        if sigmas[ax] > 1e-15:
            out = gaussian_filter1d(out, sigmas[ax], ax, orders[ax], modes[ax], cval, truncate)
            done = True
    return out if done else a.copy()


def uniform_filter(input, size=3, mode="reflect", cval=0.0, origin=0):
    """scipy/ndimage/_filters.py:1553-1622: axes with size <= 1 are skipped."""
    a = _prep(input)
    sizes = _seq(size, a.ndim)
    origins = _seq(origin, a.ndim)
    modes = _seq(mode, a.ndim)
    out = a
    done = False
    for ax in range(a.ndim):
        if sizes[ax] > 1:
            out = uniform_filter1d(out, int(sizes[ax]), ax, modes[ax], cval, origins[ax])
            done = True
    return out if done else a.copy()


def gaussian_gradient_magnitude(input, sigma, mode="reflect", cval=0.0, truncate=4.0):
    """scipy/ndimage/_filters.py:1142-1250: sqrt(sum_a gaussian(order=e_a)^2),
    squares / adds / sqrt carried out by numpy ufuncs in the ARRAY dtype."""
    a = _prep(input)
    sigmas = _seq(sigma, a.ndim)
    modes = _seq(mode, a.ndim)

    def derivative(ax):
        order = [0] * a.ndim
        order[ax] = 1
        return gaussian_filter(a, sigmas, order, modes[ax], cval, truncate)

    out = derivative(0)
    np.multiply(out, out, out)
    for ax in range(1, a.ndim):
        tmp = derivative(ax)
        np.multiply(tmp, tmp, tmp)
        out += tmp
    np.sqrt(out, out, casting="unsafe")
    return out


def gaussian_laplace(input, sigma, mode="reflect", cval=0.0, truncate=4.0):
    """scipy/ndimage/_filters.py:984-1030,1075-1139: sum_a gaussian(order=2 e_a)."""
    a = _prep(input)
    sigmas = _seq(sigma, a.ndim)
    modes = _seq(mode, a.ndim)

    def derivative2(ax):
        order = [0] * a.ndim
        order[ax] = 2
        return gaussian_filter(a, sigmas, order, modes[ax], cval, truncate)

    out = derivative2(0)
    for ax in range(1, a.ndim):
        out += derivative2(ax)
    return out


# --------------------------------------------------------------------------
# dask's map_overlap, restated (reference wrapper layer)
# --------------------------------------------------------------------------
def _chunk_bounds(n, c):
    return [(s, min(s + c, n)) for s in range(0, n, c)]


def map_overlap_emulated(func, a, chunks, depth, boundary="none", n_threads=1, **kwargs):
    """What ``image.map_overlap(func, depth=depth, boundary=boundary, **kwargs)``
    computes for the call sites dask_image/ndfilters/_gaussian.py:70-81,
    _smooth.py:28-38 and _conv.py:27-37: every chunk is extended by ``depth``
    voxels of its neighbours in every axis (``boundary="none"``: nothing is
    added past the outer faces; ``"periodic"``: the opposite end's planes are),
    ``func`` runs on the extended block and ``depth`` is trimmed again."""
    a = np.asarray(a)
    nd = a.ndim
    depth = [int(d) for d in (depth if np.iterable(depth) else [depth] * nd)]
    chunks = [int(c) for c in (chunks if np.iterable(chunks) else [chunks] * nd)]
    if boundary == "periodic":
        src = np.pad(a, [(d, d) for d in depth], mode="wrap")
        shift = depth
    else:
        src = a
        shift = [0] * nd
    out = np.empty_like(a)
    bounds = [_chunk_bounds(a.shape[d], chunks[d]) for d in range(nd)]
    jobs = [[]]
    for d in range(nd):
        jobs = [j + [b] for j in jobs for b in bounds[d]]

    def run(job):
        src_sl, trim_sl, dst_sl = [], [], []
        for d, (s, e) in enumerate(job):
            lo = s + shift[d] - depth[d]
            hi = e + shift[d] + depth[d]
            lo_c, hi_c = max(lo, 0), min(hi, src.shape[d])
            src_sl.append(slice(lo_c, hi_c))
            t0 = (s + shift[d]) - lo_c
            trim_sl.append(slice(t0, t0 + (e - s)))
            dst_sl.append(slice(s, e))
        block = func(np.ascontiguousarray(src[tuple(src_sl)]), **kwargs)
        out[tuple(dst_sl)] = block[tuple(trim_sl)]

    if n_threads > 1:
        with ThreadPoolExecutor(n_threads) as ex:
            list(ex.map(run, jobs))
    else:
        for j in jobs:
            run(j)
    return out


def gaussian_depth(sigma, truncate, ndim):
    """dask_image/ndfilters/_gaussian.py:47-55."""
    s = np.array(_seq(sigma, ndim), dtype=float)
    return tuple(int(math.ceil(v)) for v in s * truncate)
