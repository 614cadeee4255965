"""Launch single gaussian 1-D passes (strided axis 1, contiguous axis 2) for profiling."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dask_image_b200 import _lib, ndimage  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
sigma = float(sys.argv[2]) if len(sys.argv) > 2 else 2.0
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
x = torch.rand((n, n, n), dtype=torch.float32, device="cuda")
out = torch.empty_like(x)
w = ndimage._gaussian_weights(sigma, 0, 4.0)
for axis in (1, 2):
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ndimage._launch_correlate1d(x, out, tuple(x.shape), np.dtype(np.float32), axis, w, "reflect", 0.0, 0)
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(reps):
        ndimage._launch_correlate1d(x, out, tuple(x.shape), np.dtype(np.float32), axis, w, "reflect", 0.0, 0)
    ev1.record()
    torch.cuda.synchronize()
    print("axis %d %s: %.3f ms/pass" % (axis, _lib.last_kernel(), ev0.elapsed_time(ev1) / reps))
