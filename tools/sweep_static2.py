"""Static contiguous-axis kernel: CTA size / residency sweep."""
import itertools
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dask_image_b200 import ndimage  # noqa: E402
from sweep import time_pass  # noqa: E402

n = 1024
x = torch.rand((n, n, n), dtype=torch.float32, device="cuda")
out = torch.empty_like(x)
gb = 2 * 4 * x.numel() / 1e9
for sigma in (2.0, 3.0, 4.0, 5.0):
    w = ndimage._gaussian_weights(sigma, 0, 4.0)
    for thr, ctas, ncg in itertools.product((256, 128), (3, 4, 6, 8), (1, 2)):
        os.environ.update(B200_TMA_STHREADS=str(thr), B200_TMA_SCTAS=str(ctas), B200_TMA_SNCG=str(ncg))
        ms = time_pass(x, out, 2, w)
        print("sigma %g thr=%d ctas=%d ncg=%d: %.3f ms %.0f GB/s" % (sigma, thr, ctas, ncg, ms, gb / ms * 1e3), flush=True)
