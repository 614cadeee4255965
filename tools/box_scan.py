"""uint16 / uint8 box filter (uniform_filter1d) per-pass throughput by window size and axis.
    python tools/box_scan.py [planes=256]"""
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
    shape = (planes, 1024, 1024)
    vox = float(np.prod(shape))
    for code, tdt in (("uint16", torch.uint16), ("uint8", torch.uint8)):
        u = torch.randint(0, 255, shape, device="cuda", dtype=torch.int32).to(tdt)
        uo = torch.empty_like(u)
        dt = np.dtype(code)
        for size in (3, 7, 15, 31, 63, 127):
            for axis in (0, 1, 2):
                t = timeit(lambda: ndimage._launch_uniform1d(u, uo, shape, dt, axis, size, "reflect", 0.0, 0))
                gb = 2 * dt.itemsize * vox / t / 1e6
                print("%s box %3d axis %d  %-28s %7.3f ms  %6.0f GB/s  %.2f of peak" % (
                    code, size, axis, _lib.last_kernel(), t, gb, gb / PEAK), flush=True)


if __name__ == "__main__":
    main()
