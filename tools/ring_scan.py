"""Contiguous-axis static kernel: one tile per CTA vs the persistent two-stage ring."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dask_image_b200 import ndimage  # noqa: E402
from sweep import time_pass  # noqa: E402

for n in (1024, 1536):
    x = torch.rand((n, n, n), dtype=torch.float32, device="cuda")
    out = torch.empty_like(x)
    gb = 2 * 4 * x.numel() / 1e9
    for sigma in (3.0, 4.0, 5.0):
        w = ndimage._gaussian_weights(sigma, 0, 4.0)
        os.environ["B200_TMA_RING"] = "0"
        ms = time_pass(x, out, 2, w)
        print("n=%d sigma %g one-tile CTAs: %.3f ms %.0f GB/s" % (n, sigma, ms, gb / ms * 1e3), flush=True)
        os.environ["B200_TMA_RING"] = "1"
        for ctas in (2, 3, 4, 6):
            for thr in (128, 256):
                os.environ.update(B200_TMA_RING_CTAS=str(ctas), B200_TMA_STHREADS=str(thr))
                ms = time_pass(x, out, 2, w)
                print("   ring ctas=%d thr=%d: %.3f ms %.0f GB/s" % (ctas, thr, ms, gb / ms * 1e3), flush=True)
        del os.environ["B200_TMA_STHREADS"]
    del x, out
    torch.cuda.empty_cache()
