"""Per-pass throughput of the tiled kernels on rows that are NOT 16-byte multiples (plain-load staging)
next to the aligned (TMA) case: float32 weighted passes along every axis, the fused pair, the uint16 box
filter.  GB/s at the algorithmic 2 * itemsize bytes per voxel and pass.

    python tools/unaligned_scan.py [planes=256]
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dask_image_b200 import _lib, ndimage  # noqa: E402

PEAK = 6545.9


def timeit(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    e[0].record()
    for i in range(reps):
        fn()
        e[i + 1].record()
    torch.cuda.synchronize()
    return float(np.median([e[i].elapsed_time(e[i + 1]) for i in range(reps)]))


def main():
    planes = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    for side in (1024, 1023, 1034):
        shape = (planes, side, side)
        vox = float(np.prod(shape))
        x = torch.rand(shape, device="cuda")
        out = torch.empty_like(x)
        f32 = np.dtype(np.float32)
        for sigma in (2.0, 4.0):
            w = ndimage._gaussian_weights(sigma, 0, 4.0)
            for axis in (0, 1, 2):
                t = timeit(lambda: ndimage._launch_correlate1d(x, out, shape, f32, axis, w, "reflect", 0.0, 0))
                print("f32 %s sigma %.0f axis %d  %-26s %7.3f ms  %6.0f GB/s  %.2f of peak" % (
                    shape, sigma, axis, _lib.last_kernel(), t, 8 * vox / t / 1e6, 8 * vox / t / 1e6 / PEAK), flush=True)
            ps1 = ndimage._Pass("corr", 1, "reflect", weights=w)
            ps2 = ndimage._Pass("corr", 2, "reflect", weights=w)
            t = timeit(lambda: ndimage._launch_correlate1d_pair(x, out, shape, f32, ps1, ps2, 0.0))
            print("f32 %s sigma %.0f pair    %-26s %7.3f ms  %6.0f GB/s  %.2f of peak (one round trip for two passes)" % (
                shape, sigma, _lib.last_kernel(), t, 8 * vox / t / 1e6, 8 * vox / t / 1e6 / PEAK), flush=True)
        u = torch.randint(0, 65535, shape, device="cuda", dtype=torch.int32).to(torch.uint16)
        uo = torch.empty_like(u)
        u16 = np.dtype(np.uint16)
        for axis in (0, 1, 2):
            t = timeit(lambda: ndimage._launch_uniform1d(u, uo, shape, u16, axis, 7, "reflect", 0.0, 0))
            print("u16 %s box 7 axis %d     %-26s %7.3f ms  %6.0f GB/s  %.2f of peak" % (
                shape, axis, _lib.last_kernel(), t, 4 * vox / t / 1e6, 4 * vox / t / 1e6 / PEAK), flush=True)
        del x, out, u, uo
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
