"""Pass time vs volume size: does per-pass HBM throughput hold from 4 GiB to 32 GiB
volumes?  Compares every 1-D pass with a plain device copy of the same bytes."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dask_image_b200 import _lib, ndimage  # noqa: E402


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = [ev[i].elapsed_time(ev[i + 1]) for i in range(reps)]
    return min(ts), float(np.mean(ts))


def main():
    shapes = [(1024, 1024, 1024), (1536, 1536, 1536), (2048, 2048, 2048), (2048, 1024, 1024),
              (1024, 2048, 2048), (8192, 1024, 1024)]
    sigmas = [float(s) for s in sys.argv[1:]] or [2.0]
    for shape in shapes:
        x = torch.rand(shape, dtype=torch.float32, device="cuda")
        out = torch.empty_like(x)
        gb = 2 * 4 * x.numel() / 1e9
        mn, av = timed(lambda: out.copy_(x))
        print("== %s f32 (%.1f GB/pass)  torch copy: min %.3f ms %.0f GB/s, mean %.0f GB/s" % (
            shape, gb, mn, gb / mn * 1e3, gb / av * 1e3), flush=True)
        for sigma in sigmas:
            w = ndimage._gaussian_weights(sigma, 0, 4.0)
            for axis in (0, 1, 2):
                mn, av = timed(lambda: ndimage._launch_correlate1d(
                    x, out, shape, np.dtype(np.float32), axis, w, "reflect", 0.0, 0))
                print("  sigma %g axis %d: min %.3f ms %.0f GB/s, mean %.0f GB/s [%s]" % (
                    sigma, axis, mn, gb / mn * 1e3, gb / av * 1e3, _lib.last_kernel()), flush=True)
        del x, out
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
