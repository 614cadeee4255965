"""HostVolume end-to-end time vs chunk size (one GPU, 2048^3 f32 gaussian sigma=2)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dask_image_b200 as b200  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
h_in = torch.empty((n, n, n), dtype=torch.float32).pin_memory()
h_out = torch.empty_like(h_in).pin_memory()
h_in.normal_()
for mb in (128, 256, 512, 1024, 2048, 4096):
    hv = b200.HostVolume(h_in, out=h_out, chunk_bytes=mb << 20)
    b200.ndfilters.gaussian_filter(hv, 2.0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        b200.ndfilters.gaussian_filter(hv, 2.0)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 3
    print("chunk %4d MiB: %.1f ms  %.2f GVox/s  (%.1f GB/s each way)" % (mb, dt * 1e3, n ** 3 / dt / 1e9, n ** 3 * 4 / dt / 1e9),
          flush=True)
