"""Integer-dtype N-D correlate: tiled exact kernel vs the generic one (B200_FLAG_EXACT)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dask_image_b200 import _lib, ndimage  # noqa: E402
from size_scan import timed  # noqa: E402
from int_scan import TORCH  # noqa: E402


def main():
    rng = np.random.default_rng(0)
    cases = [((8192, 8192), (3, 3)), ((8192, 8192), (5, 5)), ((512, 512, 512), (3, 3, 3)), ((512, 512, 512), (5, 5, 5))]
    for code in os.environ.get("DTYPES", "u1,u2").split(","):
        dt = np.dtype(code)
        for shape, wshape in cases:
            x = torch.randint(0, 250, shape, dtype=torch.int32, device="cuda").to(TORCH[code])
            out = torch.empty_like(x)
            w = rng.standard_normal(wshape)
            for flags in (0, _lib.FLAG_EXACT):
                mn, av = timed(lambda: ndimage._launch_correlate_nd(
                    x, out, shape, dt, w, [0] * len(shape), "reflect", 0.0, flags=flags), reps=3)
                print("%s %s weights %s: %.3f ms  %.1f GVox/s [%s]" % (
                    code, shape, wshape, mn, x.numel() / mn / 1e6, _lib.last_kernel()), flush=True)
            del x, out
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
