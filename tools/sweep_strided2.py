"""Strided kernel: CTA shape / residency sweep (f32 and u16 box)."""
import itertools
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dask_image_b200 import ndimage  # noqa: E402
from sweep import time_pass  # noqa: E402

n = 1024
x = torch.rand((n, n, n), dtype=torch.float32, device="cuda")
out = torch.empty_like(x)
gb = 2 * 4 * x.numel() / 1e9
for sigma in (2.0, 3.0, 4.0):
    w = ndimage._gaussian_weights(sigma, 0, 4.0)
    for axis in (0, 1):
        for ty, ctas in itertools.product((8, 4), (0, 3, 4, 6, 8)):
            os.environ.update(B200_TMA_TY=str(ty), B200_TMA_CTAS=str(ctas))
            ms = time_pass(x, out, axis, w)
            print("f32 sigma %g axis %d ty=%d ctas=%d: %.3f ms %.0f GB/s" % (sigma, axis, ty, ctas, ms, gb / ms * 1e3),
                  flush=True)
del x, out
u = torch.randint(0, 65536, (n, n, n), dtype=torch.int32, device="cuda").to(torch.uint16)
uo = torch.empty_like(u)
for axis in (0, 1):
    for ty, ctas in itertools.product((8, 4), (0, 3, 4, 6, 8)):
        os.environ.update(B200_TMA_TY=str(ty), B200_TMA_CTAS=str(ctas))
        ndimage._launch_uniform1d(u, uo, (n, n, n), np.dtype(np.uint16), axis, 7, "reflect", 0.0, 0)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
        ev[0].record()
        for i in range(5):
            ndimage._launch_uniform1d(u, uo, (n, n, n), np.dtype(np.uint16), axis, 7, "reflect", 0.0, 0)
            ev[i + 1].record()
        torch.cuda.synchronize()
        ms = min(ev[i].elapsed_time(ev[i + 1]) for i in range(5))
        print("u16 box7 axis %d ty=%d ctas=%d: %.3f ms %.0f GB/s" % (axis, ty, ctas, ms, gb / 2 / ms * 1e3), flush=True)
