"""Generate the golden vectors under tests/golden/ by running the reference's
per-chunk arithmetic (scipy.ndimage, as called from
dask_image/dispatch/_dispatch_ndfilters.py:49,65,129,145,161,273) in THIS
container.  The GPU box has no /root/reference, so the vectors travel as
fixtures; this script is committed so they can be regenerated:

    python tests/golden/make_golden.py

The stored outputs are also what the REFERENCE ITSELF computes (its wrappers + numpy registrations, imported
unmodified with a stand-in for dask.utils.Dispatch): tests/test_reference_wrappers_cpu.py::
test_goldens_are_the_reference_outputs recomputes every public-function case that way and requires bit equality.

Parametrisations follow the reference's own tests
(tests/test_dask_image/test_ndfilters/test__gaussian.py:125-196,
test__smooth.py:80-99, test__conv.py:117-191) plus a dtype sweep.
"""
import json
import os

import numpy as np
import scipy
import scipy.ndimage as ndi

HERE = os.path.dirname(os.path.abspath(__file__))

cases = []   # (name, func, input, kwargs)


def add(name, func, a, **kw):
    cases.append((name, func, np.ascontiguousarray(a), kw))


def main():
    rng = np.random.default_rng(20261019)
    a_small = np.arange(140.0).reshape(10, 14)
    a_int = np.arange(140).reshape(10, 14)
    a_big = np.arange(float(100 * 110)).reshape(100, 110)
    a_mid = np.arange(float(40 * 44)).reshape(40, 44)

    # gaussian family (test__gaussian.py:125-196)
    for sigma, truncate in [(1.0, 2.0), (1.0, 4.0), (2.0, 2.0), (2.0, 4.0), ((1.0, 2.0), 4.0)]:
        for fn in ("gaussian_filter", "gaussian_gradient_magnitude", "gaussian_laplace"):
            add("ref_%s_s%s_t%s" % (fn, sigma, truncate), fn, a_big, sigma=sigma, truncate=truncate)
    for sigma, truncate in [(0.0, 0.0), (1.0, 0.0), (0.0, 1.0), (1.0, 2.0), (2.0, 4.0),
                            ((1.0, 2.0), 4.0)]:
        for order in [0, 1, 2, 3, (0, 1), (2, 3)]:
            add("ref_gauss_order%s_s%s_t%s" % (order, sigma, truncate), "gaussian_filter", a_mid,
                sigma=sigma, order=order, truncate=truncate)
    # uniform (test__smooth.py:80-99)
    for size, origin in [(1, 0), (2, 0), (3, 0), (3, 1), (3, (1, 0)), ((1, 2), 0), ((3, 2), (1, 0))]:
        add("ref_uniform_%s_%s" % (size, origin), "uniform_filter", a_big, size=size, origin=origin)
    # conv (test__conv.py:117-191)
    disc = (np.mgrid[-2:3, -2:3] ** 2).sum(axis=0) < 2.5 ** 2
    wlist = [(np.ones((1, 1)), 0), (np.ones((2, 2)), 0), (np.ones((2, 3)), 0), (np.ones((2, 3)), (0, 1)),
             (np.ones((2, 3)), (0, -1)), (disc, 0), (disc, (1, 2)), (disc, (-1, -2)),
             (np.ones((5, 5)), 0), (np.ones((7, 7)), 0), (np.ones((8, 8)), 0),
             (np.ones((10, 10)), 0), (np.ones((5, 5)), 2), (np.ones((5, 5)), -2)]
    for i, (w, origin) in enumerate(wlist):
        for fn in ("convolve", "correlate"):
            add("ref_%s_w%d" % (fn, i), fn, a_small, weights=w, origin=origin)
    for mode in ["reflect", "wrap", "nearest", "constant", "mirror"]:
        for w in (np.ones((1, 5)), np.ones((5, 1))):
            for fn in ("convolve", "correlate"):
                add("ref_%s_modes_%s_%s" % (fn, mode, w.shape), fn, a_int, weights=w, mode=mode)
    # dtype sweep, all modes, 3-D
    dts = ["uint8", "int8", "uint16", "int16", "uint32", "int32", "uint64", "int64", "float32", "float64"]
    for dt in dts:
        if dt.startswith("f"):
            a = (rng.random((9, 12, 16)) * 200 - 50).astype(dt)
        else:
            info = np.iinfo(dt)
            a = rng.integers(max(info.min, -3000), min(info.max, 3000), (9, 12, 16), endpoint=True).astype(dt)
        for mode in ["reflect", "wrap", "nearest", "constant", "mirror"]:
            add("dt_gauss_%s_%s" % (dt, mode), "gaussian_filter", a, sigma=(1.0, 1.5, 2.0), mode=mode, cval=3.5)
            add("dt_gauss_d_%s_%s" % (dt, mode), "gaussian_filter", a, sigma=1.2, order=(0, 1, 2), mode=mode,
                cval=3.5)
            add("dt_uniform_%s_%s" % (dt, mode), "uniform_filter", a, size=(3, 4, 7), origin=(0, -1, 2),
                mode=mode, cval=3.5)
            w = rng.random((3, 2, 4)) - 0.4
            w[1, 0, 2] = 0.0
            add("dt_corr_%s_%s" % (dt, mode), "correlate", a, weights=w, origin=(1, -1, 0), mode=mode, cval=3.5)
            add("dt_conv_%s_%s" % (dt, mode), "convolve", a, weights=w, origin=(-1, 0, 1), mode=mode, cval=3.5)
        if dt in ("uint8", "uint16", "int32", "float32", "float64"):
            add("dt_ggm_%s" % dt, "gaussian_gradient_magnitude", a, sigma=1.0)
            add("dt_glap_%s" % dt, "gaussian_laplace", a, sigma=1.0)
    # 1-D primitives incl. radius >= axis length
    x = rng.random(7)
    for mode in ["reflect", "wrap", "nearest", "constant", "mirror"]:
        add("c1d_long_%s" % mode, "correlate1d", x, weights=rng.random(23), mode=mode, cval=-1.0)
        add("c1d_even_%s" % mode, "correlate1d", x, weights=rng.random(4), origin=-2, mode=mode, cval=-1.0)
        add("u1d_long_%s" % mode, "uniform_filter1d", x, size=19, origin=3, mode=mode, cval=-1.0)

    arrays, spec, seen = {}, [], {}
    for i, (name, fn, a, kw) in enumerate(cases):
        expected = getattr(ndi, fn)(a, **kw)
        key = (a.dtype.str, a.shape, a.tobytes())
        if key not in seen:                     # inputs are stored once
            seen[key] = "in_%d" % i
            arrays[seen[key]] = a
        arrays["out_%d" % i] = expected
        kws = {}
        for k, v in kw.items():
            if isinstance(v, np.ndarray):
                arrays["kw_%d_%s" % (i, k)] = v
                kws[k] = "@array"
            else:
                kws[k] = v
        spec.append({"name": name, "func": fn, "input": seen[key], "kwargs": kws})
    np.savez_compressed(os.path.join(HERE, "ndfilters_golden.npz"), **arrays)
    with open(os.path.join(HERE, "ndfilters_golden.json"), "w") as f:
        json.dump({"scipy": scipy.__version__, "numpy": np.__version__, "cases": spec}, f, indent=0)
    print("wrote %d cases" % len(cases))


if __name__ == "__main__":
    main()
