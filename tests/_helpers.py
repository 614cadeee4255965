"""Shared helpers of the test-suite: golden fixtures, tolerances, oracle access."""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN_NPZ = os.path.join(HERE, "golden", "ndfilters_golden.npz")
GOLDEN_JSON = os.path.join(HERE, "golden", "ndfilters_golden.json")

# north_star tolerances (BASELINE.json): fp32 <= 1e-5 relative with fp32
# accumulation, fp64 <= 1e-12, integers bit-exact.  Derivative / laplace
# outputs cross zero, so "relative" is taken against the output's magnitude as
# well as elementwise:
#     |got - ref| <= rtol * |ref| + atol_scale * max|ref|      (atol_scale = rtol / 5, below)
# Where the caller passes the filter's input (``data=``) the scale is
# max(max|ref|, max|data|): a derivative kernel applied to a constant image is
# EXACTLY zero in scipy (it folds x[-j] - x[j] first) while any direct-form
# kernel leaves rounding noise of order eps * sum|w| * |data| -- an error
# relative to the data, not to an all-zero reference.
RTOL = {np.dtype(np.float32): 1e-5, np.dtype(np.float64): 1e-12}
# The absolute floor (the term relative to the output's / data's magnitude) is five times tighter
# than the elementwise one: observed errors are ~3e-7 (fp32) / ~3e-16 (fp64) of the scale, so the
# bound still catches a 10x regression.  Achieved maxima per test are collected in ACHIEVED and
# written to gpurun_out/parity_errors.json at the end of a GPU session (conftest.py).
ATOL_SCALE = {np.dtype(np.float32): 2e-6, np.dtype(np.float64): 2e-13}
ACHIEVED = {}

_golden = None


def golden_cases():
    with open(GOLDEN_JSON) as f:
        return json.load(f)["cases"]


def golden_arrays():
    global _golden
    if _golden is None:
        _golden = dict(np.load(GOLDEN_NPZ))
    return _golden


def golden_case(i):
    spec = golden_cases()[i]
    arrs = golden_arrays()
    kwargs = {}
    for k, v in spec["kwargs"].items():
        if v == "@array":
            kwargs[k] = arrs["kw_%d_%s" % (i, k)]
        elif isinstance(v, list):
            kwargs[k] = tuple(v)
        else:
            kwargs[k] = v
    return spec["func"], arrs[spec["input"]], kwargs, arrs["out_%d" % i]


def assert_parity(got, ref, exact=False, what="", data=None):
    got = np.asarray(got)
    ref = np.asarray(ref)
    assert got.shape == ref.shape, "%s shape %s != %s" % (what, got.shape, ref.shape)
    assert got.dtype == ref.dtype, "%s dtype %s != %s" % (what, got.dtype, ref.dtype)
    if exact or ref.dtype.kind != "f":
        if not np.array_equal(got, ref, equal_nan=True):
            bad = np.argwhere(got != ref)
            raise AssertionError("%s: %d / %d elements differ (first at %s: got %r, ref %r)" % (
                what, len(bad), ref.size, tuple(bad[0]), got[tuple(bad[0])], ref[tuple(bad[0])]))
        return
    rtol = RTOL[ref.dtype]
    scale = float(np.max(np.abs(ref))) if ref.size else 0.0
    if data is not None and np.size(data):
        scale = max(scale, float(np.max(np.abs(np.asarray(data, dtype=np.float64)))))
    err = np.abs(got.astype(np.float64) - ref.astype(np.float64))
    bound = rtol * np.abs(ref.astype(np.float64)) + ATOL_SCALE[ref.dtype] * scale
    if err.size and scale > 0:
        import os

        key = "%s:%s" % (os.environ.get("PYTEST_CURRENT_TEST", "?").split("::")[-1].split("[")[0].split(" ")[0],
                         ref.dtype.name)
        ACHIEVED[key] = max(ACHIEVED.get(key, 0.0), float(np.nanmax(err)) / scale)
    if not np.all(err <= bound):
        i = np.unravel_index(np.argmax(err - bound), err.shape)
        raise AssertionError("%s: |got-ref|=%g exceeds %g at %s (got %r, ref %r, rtol %g)" % (
            what, err[i], bound[i], i, got[i], ref[i], rtol))
