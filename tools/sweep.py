"""Time single 1-D passes under different kernel tuning knobs (env vars read at launch)."""
import itertools
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dask_image_b200 import _lib, ndimage  # noqa: E402


def time_pass(x, out, axis, w, reps=5):
    shape, dt = tuple(x.shape), np.dtype(np.float32)
    ndimage._launch_correlate1d(x, out, shape, dt, axis, w, "reflect", 0.0, 0)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        ndimage._launch_correlate1d(x, out, shape, dt, axis, w, "reflect", 0.0, 0)
        ev[i + 1].record()
    torch.cuda.synchronize()
    return min(ev[i].elapsed_time(ev[i + 1]) for i in range(reps))


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    x = torch.rand((n, n, n), dtype=torch.float32, device="cuda")
    out = torch.empty_like(x)
    gb = 2 * 4 * x.numel() / 1e9
    for k in list(os.environ):
        if k.startswith("B200_TMA_"):
            del os.environ[k]
    for sigma in (1.0, 2.0, 4.0, 8.0):
        w = ndimage._gaussian_weights(sigma, 0, 4.0)
        print("== sigma %g (%d taps), %d^3 f32, %.2f GB/pass" % (sigma, len(w), n, gb), flush=True)
        for axis in (0, 1):
            for R, NB in itertools.product((4,), (0,)):
                os.environ.update(B200_TMA_R=str(R), B200_TMA_NB=str(NB))
                try:
                    ms = time_pass(x, out, axis, w)
                    print("strided axis%d R=%d NB=%d : %.3f ms  %.0f GB/s  [%s]" % (
                        axis, R, NB, ms, gb / ms * 1e3, _lib.last_kernel()), flush=True)
                except Exception as e:
                    print("strided axis%d R=%d NB=%d : ERR %s" % (axis, R, NB, e))
        del os.environ["B200_TMA_R"], os.environ["B200_TMA_NB"]
        for NCG, TR in itertools.product((2, 3, 4, 5, 6), (64, 96, 128, 160, 256)):
            os.environ.update(B200_TMA_NCG=str(NCG), B200_TMA_TR=str(TR))
            ms = time_pass(x, out, 2, w)
            print("contig NCG=%d TR=%d : %.3f ms  %.0f GB/s [%s]" % (
                NCG, TR, ms, gb / ms * 1e3, _lib.last_kernel()), flush=True)
        del os.environ["B200_TMA_NCG"], os.environ["B200_TMA_TR"]


if __name__ == "__main__":
    main()
