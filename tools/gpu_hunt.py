"""Time-boxed random parity hunt on the GPU: the public per-chunk functions against scipy on volumes big enough to
cross tile boundaries in every direction (several tiles of rows, columns and planes; aligned and odd widths).

    python tools/gpu_hunt.py [seconds=20] [seed=0]

Complements tests/test_gpu_fuzz.py (small shapes, fixed examples): here the shapes are 0.2-3 MVoxels, the seed is a
command-line argument and the run lasts as long as it is given.  Prints one line per failure with everything needed to
reproduce it, then a summary line; exit code 1 if anything differed.
"""
import os
import sys
import time

import numpy as np
import scipy.ndimage as ndi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

MODES = ["reflect", "constant", "nearest", "mirror", "wrap"]
DTYPES = ["float32", "float32", "uint16", "uint8", "int16", "float64"]


def draw_shape(rng):
    nd = int(rng.choice([2, 3, 3, 3]))
    # last axis: around / across the 256-column (float) and 512-byte tiles, aligned and not
    last = int(rng.choice([256, 264, 272, 512, 520, 1024, 250, 255, 257, 263, 300, 511, 513, 96, 40]))
    mid = int(rng.choice([31, 32, 33, 40, 63, 64, 65, 70, 96, 100, 129]))
    if nd == 2:
        return (int(rng.choice([33, 64, 100, 257, 300, 700])), last)
    lead = int(rng.integers(1, 12))
    while lead * mid * last > 3_000_000:
        lead = max(1, lead // 2)
        if lead == 1:
            break
    return (lead, mid, last)


def data_for(rng, shape, dt):
    dt = np.dtype(dt)
    if dt.kind == "f":
        return (rng.random(shape) * 200 - 60).astype(dt)
    ii = np.iinfo(dt)
    return rng.integers(ii.min, ii.max, shape, endpoint=True).astype(dt)


def draw_case(rng):
    shape = draw_shape(rng)
    nd = len(shape)
    dt = str(rng.choice(DTYPES))
    mode = str(rng.choice(MODES))
    cval = float(rng.choice([0.0, 3.0, 100.0]))
    kind = str(rng.choice(["gaussian_iso", "gaussian_iso", "gaussian_any", "uniform_iso", "uniform_iso", "uniform_any",
                           "ggm", "laplace", "correlate1d", "correlate_nd", "minmax", "sobel"]))
    if kind == "gaussian_iso":
        s = float(rng.choice([0.5, 1.0, 1.5, 2.0, 2.5, 3.0, 4.0, 5.0]))
        return ("gaussian_filter", shape, dt, (s,), dict(mode=mode, cval=cval))
    if kind == "gaussian_any":
        s = tuple(float(rng.choice([0.0, 0.7, 1.0, 2.0, 3.3])) for _ in range(nd))
        o = tuple(int(rng.integers(0, 3)) for _ in range(nd))
        return ("gaussian_filter", shape, dt, (s,), dict(order=o, mode=mode, cval=cval,
                                                         truncate=float(rng.choice([4.0, 3.0, 2.0]))))
    if kind == "uniform_iso":
        return ("uniform_filter", shape, dt, (int(rng.choice([3, 5, 7, 9, 11, 15, 17, 19, 31])),),
                dict(mode=mode, cval=cval))
    if kind == "uniform_any":
        size = tuple(int(rng.integers(1, 12)) for _ in range(nd))
        origin = tuple(int(rng.integers(-(s // 2), (s - 1) // 2 + 1)) for s in size)
        return ("uniform_filter", shape, dt, (size,), dict(mode=mode, cval=cval, origin=origin))
    if kind in ("ggm", "laplace"):
        s = float(rng.choice([0.8, 1.0, 2.0, 3.0]))
        return ("gaussian_gradient_magnitude" if kind == "ggm" else "gaussian_laplace", shape, dt, (s,),
                dict(mode=mode, cval=cval))
    if kind == "correlate1d":
        nw = int(rng.integers(1, 50))
        w = rng.random(nw) - 0.4
        if rng.random() < 0.4:
            w = w + w[::-1]
        origin = int(rng.integers(-(nw // 2), (nw - 1) // 2 + 1))
        return ("correlate1d", shape, dt, (w,), dict(axis=int(rng.integers(0, nd)), mode=mode, cval=cval,
                                                      origin=origin))
    if kind == "correlate_nd":
        wshape = tuple(int(rng.integers(1, 6)) for _ in range(nd))
        w = rng.random(wshape) - 0.4
        if rng.random() < 0.3:
            w[w < 0] = 0.0
        origin = tuple(int(rng.integers(-(s // 2), (s - 1) // 2 + 1)) for s in wshape)
        return (str(rng.choice(["correlate", "convolve"])), shape, dt, (w,),
                dict(mode=mode, cval=cval, origin=origin) if rng.random() < 0.5 else dict(mode=mode, cval=cval))
    if kind == "minmax":
        size = tuple(int(rng.integers(1, 9)) for _ in range(nd))
        return (str(rng.choice(["minimum_filter", "maximum_filter"])), shape, dt, (), dict(size=size, mode=mode,
                                                                                          cval=cval))
    return (str(rng.choice(["sobel", "prewitt"])), shape, dt, (), dict(axis=int(rng.integers(0, nd)), mode=mode,
                                                                      cval=cval))


def main():
    seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 20.0
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    import dask_image_b200 as b200
    from _helpers import assert_parity
    from dask_image_b200 import _lib

    rng = np.random.default_rng(seed)
    t0 = time.time()
    n = bad = 0
    kernels = {}
    voxels = 0
    while time.time() - t0 < seconds:
        name, shape, dt, args, kw = draw_case(rng)
        a = data_for(rng, shape, dt)
        with np.errstate(all="ignore"):
            ref = getattr(ndi, name)(a, *args, **kw)
            got = getattr(b200.ndimage, name)(b200.B200Array.from_numpy(a), *args, **kw).to_numpy()
        k = _lib.last_kernel()
        kernels[k] = kernels.get(k, 0) + 1
        n += 1
        voxels += a.size
        try:
            # the float box filter: scipy's running sum rounds per line start, the kernels sum directly
            assert_parity(got, ref, what=name, data=a)
        except AssertionError as e:
            bad += 1
            print("FAIL seed=%d case=%d %s shape=%s dtype=%s kw=%r args=%s kernel=%s: %s" % (
                seed, n, name, shape, dt, {k_: v for k_, v in kw.items()},
                [x if np.isscalar(x) or isinstance(x, tuple) else getattr(x, "shape", x) for x in args], k,
                str(e)[:200]), flush=True)
    print("hunt: seed %d, %d cases, %.1f MVoxels, %d failures, %.1f s; last kernels: %s" % (
        seed, n, voxels / 1e6, bad, time.time() - t0, dict(sorted(kernels.items()))))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
