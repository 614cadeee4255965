"""Launch each hot kernel a few times on a 1024^3 volume (for ncu captures).
Usage: prof_kernels.py [f32|u16|nd] [sigma] [reps]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dask_image_b200 import _lib, ndimage  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "f32"
sigma = float(sys.argv[2]) if len(sys.argv) > 2 else 2.0
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
n = int(os.environ.get("B200_PROF_N", "1024"))


def timed(fn, label):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print("%s [%s]: %.3f ms" % (label, _lib.last_kernel(), e0.elapsed_time(e1) / reps), flush=True)


if what == "f32":
    x = torch.rand((n, n, n), dtype=torch.float32, device="cuda")
    out = torch.empty_like(x)
    w = ndimage._gaussian_weights(sigma, 0, 4.0)
    for axis in (0, 1, 2):
        timed(lambda: ndimage._launch_correlate1d(x, out, tuple(x.shape), np.dtype(np.float32), axis, w,
                                                  "reflect", 0.0, 0), "gauss sigma %g axis %d" % (sigma, axis))
elif what == "f64":
    m = 768
    x = torch.rand((m, m, m), dtype=torch.float64, device="cuda")
    out = torch.empty_like(x)
    w = ndimage._gaussian_weights(sigma, 0, 4.0)
    gb = 2 * 8 * x.numel() / 1e9
    for axis in (0, 1, 2):
        timed(lambda: ndimage._launch_correlate1d(x, out, tuple(x.shape), np.dtype(np.float64), axis, w,
                                                  "reflect", 0.0, 0), "f64 %d^3 (%.1f GB/pass) gauss sigma %g axis %d" % (m, gb, sigma, axis))
elif what == "u16":
    x = torch.randint(0, 65536, (n, n, n), dtype=torch.int32, device="cuda").to(torch.uint16)
    out = torch.empty_like(x)
    for axis in (0, 1, 2):
        timed(lambda: ndimage._launch_uniform1d(x, out, tuple(x.shape), np.dtype(np.uint16), axis, 7, "reflect",
                                                0.0, 0), "uniform7 axis %d" % axis)
else:
    m = int(os.environ.get("B200_PROF_ND", "512"))
    x = torch.rand((m, m, m), dtype=torch.float32, device="cuda")
    out = torch.empty_like(x)
    k = int(sigma) if sigma >= 1 else 9
    w = np.random.default_rng(0).random((k, k, k))
    timed(lambda: ndimage._launch_correlate_nd(x, out, tuple(x.shape), np.dtype(np.float32), w, (0, 0, 0), "wrap",
                                               0.0), "correlate %d^3 on %d^3" % (k, m))
