"""Large one-off random comparison of the host-emulated integer kernels (tests/native/int_emul.cu, both walk
widths) with scipy.ndimage: python tools/emul_fuzz.py [seed]  (CASES=1500).  Needs the emulation library
built (run tests/test_int_emul.py once)."""
import sys, os, ctypes, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import scipy.ndimage as ndi
import test_int_emul as T
lib = ctypes.CDLL(T.OUT)
lib.emul_correlate1d_int.restype = ctypes.c_int
lib.emul_correlate_nd_int.restype = ctypes.c_int
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
n1 = nn = 0
sizes = [1, 2, 3, 4, 5, 7, 8, 9, 15, 16, 17, 31, 32, 33, 40, 64, 67, 96, 127, 128, 129, 130, 160, 257]
for it in range(int(os.environ.get("CASES", "1500"))):
    nd = int(rng.integers(1, 4))
    shape = tuple(int(rng.choice(sizes[: (len(sizes) if nd < 3 else 14)])) for _ in range(nd))
    dt = np.dtype(rng.choice(T.DTYPES))
    mode = str(rng.choice(T.MODES))
    ii = np.iinfo(dt)
    cval = float(rng.integers(max(ii.min, -100), min(ii.max, 100))) if mode == "constant" else 0.0
    a = rng.integers(ii.min, ii.max + 1, size=shape, dtype=np.int64).astype(dt)
    if rng.random() < 0.2:
        a[...] = ii.max if rng.random() < 0.5 else ii.min      # saturated: largest pair sums
    axis = int(rng.integers(0, nd))
    nw = int(rng.integers(1, 60))
    origin = int(rng.integers(-(nw // 2), (nw - 1) // 2 + 1))
    w = rng.standard_normal(nw) * float(rng.choice([1e-3, 1.0, 50.0]))
    if nw % 2 and rng.random() < 0.6:
        w = (w + w[::-1]) / 2 if rng.random() < 0.5 else (w - w[::-1]) / 2
    os.environ["B200_INT_NQ"] = "8" if rng.random() < 0.5 else "4"
    got, info = T.run_emul(lib, a, w, axis, mode, cval, origin)
    assert got is not None
    ref = ndi.correlate1d(a, w, axis=axis, mode=mode, cval=cval, origin=origin)
    assert np.array_equal(got, ref), (it, shape, dt, axis, mode, nw, origin, info)
    n1 += 1
    if nd >= 2:
        wshape = tuple(int(rng.integers(1, 6)) for _ in range(nd))
        wn = rng.standard_normal(wshape)
        origins = [int(rng.integers(-(s // 2), (s - 1) // 2 + 1)) for s in wshape]
        got, info = T.run_emul_nd(lib, a, wn, mode, cval, origins)
        assert got is not None
        assert np.array_equal(got, ndi.correlate(a, wn, mode=mode, cval=cval, origin=origins)), (it, shape, dt, wshape, mode, origins)
        nn += 1
print("ok", n1, "1-D cases,", nn, "N-D cases")
