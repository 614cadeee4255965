"""Integer-dtype correlate1d passes (gaussian weights): time per pass and HBM throughput."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dask_image_b200 import _lib, ndimage  # noqa: E402
from size_scan import timed  # noqa: E402

TORCH = {"u1": torch.uint8, "u2": torch.uint16, "i2": torch.int16, "i4": torch.int32}


def main():
    n = int(os.environ.get("N", "1024"))
    shape = (n, n, n)
    sigmas = [float(s) for s in sys.argv[1:]] or [1.0, 2.0, 4.0]
    for code in os.environ.get("DTYPES", "u1,u2").split(","):
        dt = np.dtype(code)
        x = torch.randint(0, 250, shape, dtype=torch.int32, device="cuda").to(TORCH[code])
        out = torch.empty_like(x)
        gb = 2 * dt.itemsize * x.numel() / 1e9
        for sigma in sigmas:
            for order in (0, 1, -1):
                w = ndimage._gaussian_weights(sigma, max(order, 0), 4.0)
                if order < 0:   # neither symmetric nor antisymmetric: the single-tap loop
                    w = w + np.linspace(0.0, 0.01, w.size)
                for axis in (0, 1, 2):
                    mn, av = timed(lambda: ndimage._launch_correlate1d(
                        x, out, shape, dt, axis, w, "reflect", 0.0, 0), reps=3)
                    print("%s sigma %g order %d axis %d: %.3f ms %.0f GB/s  %.0f GVox/s [%s]" % (
                        code, sigma, order, axis, mn, gb / mn * 1e3, x.numel() / mn / 1e6, _lib.last_kernel()),
                        flush=True)
        del x, out
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
