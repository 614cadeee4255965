"""Tile-geometry sweep of the fused pass pair (corr2d_fused.cu) on the visible B200.

Needs a library built with NVCC_EXTRA=-DB200_FUSED_SWEEP (the variants are
compiled for the sigma = 2 / 3 / 4 pairs only).  For every sigma and variant:
time the pair launch on a resident volume with CUDA events, check the result
bit for bit against the two one-pass launches, print ms and GB/s at the
algorithmic 8 B/voxel next to the two-launch time.

    python tools/sweep_fused.py [N=1024] [sigmas...]
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dask_image_b200 import _lib, ndimage  # noqa: E402

VARIANTS = {0: "default (32x8w RA8 alias)", 1: "32x8w RA8 noalias minb2", 2: "32x8w RA8 alias minb2",
            3: "32x8w RA8 alias minb4", 4: "16x4w RA8 alias minb6", 6: "64x16w RA8 alias minb1",
            7: "24x6w RA8 alias minb5", 8: "24x6w RA8 alias minb4"}
if os.environ.get("SWEEP_VARIANTS"):
    VARIANTS = {int(v): VARIANTS[int(v)] for v in os.environ["SWEEP_VARIANTS"].split(",")}


def timeit(fn, reps):
    fn()
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    e[0].record()
    for i in range(reps):
        fn()
        e[i + 1].record()
    torch.cuda.synchronize()
    ts = [e[i].elapsed_time(e[i + 1]) for i in range(reps)]
    return float(np.median(ts)), float(np.min(ts))


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    sigmas = [float(s) for s in sys.argv[2:]] or [2.0, 3.0, 4.0]
    shape = (n, n, n) if n <= 1280 else (n // 4, n, n)
    vox = int(np.prod(shape))
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.rand(shape, device="cuda", generator=g)
    tmp, ref, out = torch.empty_like(x), torch.empty_like(x), torch.empty_like(x)
    dt = np.dtype(np.float32)
    reps = int(os.environ.get("REPS", "10"))
    for sigma in sigmas:
        w = ndimage._gaussian_weights(sigma, 0, 4.0)
        ps1 = ndimage._Pass("corr", 1, "reflect", weights=w)
        ps2 = ndimage._Pass("corr", 2, "reflect", weights=w)

        def two():
            ndimage._launch_correlate1d(x, tmp, shape, dt, 1, w, "reflect", 0.0, 0)
            ndimage._launch_correlate1d(tmp, ref, shape, dt, 2, w, "reflect", 0.0, 0)

        t2, t2min = timeit(two, reps)
        print("sigma %.1f shape %s taps %d: two launches %.3f ms (min %.3f) = %.0f GB/s at 16 B/voxel" % (
            sigma, shape, len(w), t2, t2min, 16 * vox / t2 / 1e6), flush=True)
        for v, name in VARIANTS.items():
            os.environ["B200_FUSED_VARIANT"] = str(v)
            try:
                out.zero_()
                t, tmin = timeit(lambda: ndimage._launch_correlate1d_pair(x, out, shape, dt, ps1, ps2, 0.0), reps)
            except Exception as exc:
                print("  v%d %-30s FAILED %s" % (v, name, str(exc)[:100]), flush=True)
                continue
            same = bool(torch.equal(out, ref))
            print("  v%d %-30s %.3f ms (min %.3f) = %.0f GB/s at 8 B/voxel, %.2fx vs two, bit-identical=%s [%s]" % (
                v, name, t, tmin, 8 * vox / t / 1e6, t2 / t, same, _lib.last_kernel()), flush=True)
        os.environ.pop("B200_FUSED_VARIANT", None)


if __name__ == "__main__":
    main()
