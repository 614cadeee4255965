"""Sweep tile geometry of the static contiguous-axis kernel."""
import itertools
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dask_image_b200 import _lib, ndimage  # noqa: E402
from sweep import time_pass  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
x = torch.rand((n, n, n), dtype=torch.float32, device="cuda")
out = torch.empty_like(x)
gb = 2 * 4 * x.numel() / 1e9
for sigma in (1.0, 2.0, 3.0, 4.0, 5.0):
    w = ndimage._gaussian_weights(sigma, 0, 4.0)
    os.environ["B200_TMA_STATIC"] = "0"
    ms = time_pass(x, out, 2, w)
    print("sigma %g (%d taps) run-time engine: %.3f ms %.0f GB/s" % (sigma, len(w), ms, gb / ms * 1e3), flush=True)
    os.environ["B200_TMA_STATIC"] = "1"
    for ncg, tr in itertools.product((1, 2), (0, 32, 48, 64, 80, 96, 128)):
        os.environ.update(B200_TMA_SNCG=str(ncg), B200_TMA_STR=str(tr))
        try:
            ms = time_pass(x, out, 2, w)
            print("  static ncg=%d tr=%3d: %.3f ms %.0f GB/s" % (ncg, tr, ms, gb / ms * 1e3), flush=True)
        except Exception as e:
            print("  static ncg=%d tr=%d: ERR %s" % (ncg, tr, e))
    del os.environ["B200_TMA_SNCG"], os.environ["B200_TMA_STR"]
