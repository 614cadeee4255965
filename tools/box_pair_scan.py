"""uint16 / uint8 box filter: the fused pair (last two axes, one launch) against the two one-pass kernels.
    python tools/box_pair_scan.py [planes=512]         (planes x 2048 x 2048)"""
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
    planes = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    shape = (planes, 2048, 2048)
    vox = float(np.prod(shape))
    for code, tdt in (("uint16", torch.uint16), ("uint8", torch.uint8)):
        dt = np.dtype(code)
        u = torch.randint(0, 256 ** dt.itemsize, shape, device="cuda", dtype=torch.int32).to(tdt)
        mid = torch.empty_like(u)
        uo = torch.empty_like(u)
        sizes = (3, 7, 9, 17) if code == "uint16" else (7,)
        for size in sizes:
            ps = [ndimage._Pass("uniform", ax, "reflect", size=size) for ax in (1, 2)]

            def two():
                ndimage._launch_uniform1d(u, mid, shape, dt, 1, size, "reflect", 0.0, 0)
                ndimage._launch_uniform1d(mid, uo, shape, dt, 2, size, "reflect", 0.0, 0)

            def pair():
                ndimage._launch_correlate1d_pair(u, uo, shape, dt, ps[0], ps[1], 0.0)

            t2 = timeit(two)
            want = uo.clone()
            rows = [("two launches", t2)]
            for slide in ("1", "0"):
                if slide == "0" and (size != 7 or code != "uint16"):     # the direct form exists for uint16, size 7
                    continue
                os.environ["B200_BOX2D_SLIDE"] = slide
                uo.zero_()
                tp = timeit(pair)
                assert torch.equal(uo.view(torch.uint8), want.view(torch.uint8)), "fused != two passes"
                rows.append(("pair %s (%s)" % ("running sums" if slide == "1" else "direct", _lib.last_kernel()), tp))
            os.environ["B200_BOX2D_SLIDE"] = "1"
            for name, t in rows:
                gb = 2 * dt.itemsize * vox / t / 1e6        # one read + one write of the volume
                print("%s size %2d  %-50s %7.3f ms  %6.0f GB/s moved (2 x %d B/voxel)  %.2f of peak as ONE pass" % (
                    code, size, name, t, gb, dt.itemsize, gb / PEAK), flush=True)


if __name__ == "__main__":
    main()
