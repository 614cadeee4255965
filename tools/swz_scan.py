"""Axis-0 pass time vs tile-order group width (B200_TMA_SWZ)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dask_image_b200 import _lib, ndimage  # noqa: E402
from sweep import time_pass  # noqa: E402

for shape in [(1024, 1024, 1024), (2048, 2048, 2048), (256, 2048, 2048)]:
    x = torch.rand(shape, dtype=torch.float32, device="cuda")
    out = torch.empty_like(x)
    gb = 2 * 4 * x.numel() / 1e9
    for sigma in (2.0, 4.0):
        w = ndimage._gaussian_weights(sigma, 0, 4.0)
        ms1 = time_pass(x, out, 1, w)
        print("%s sigma %g axis1: %.3f ms %.0f GB/s" % (shape, sigma, ms1, gb / ms1 * 1e3), flush=True)
        for swz in (1, 32, 64, 128, 256, 512, 1024, 4096):
            os.environ["B200_TMA_SWZ"] = str(swz)
            ms = time_pass(x, out, 0, w)
            print("   axis0 swz=%4d: %.3f ms %.0f GB/s" % (swz, ms, gb / ms * 1e3), flush=True)
        del os.environ["B200_TMA_SWZ"]
    del x, out
    torch.cuda.empty_cache()
