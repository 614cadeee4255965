"""A few launches of the fused uint16 box pair (size 7) on 512 x 2048 x 2048, for an ncu capture."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dask_image_b200 import _lib, ndimage  # noqa: E402

size = int(sys.argv[1]) if len(sys.argv) > 1 else 7
shape = (512, 2048, 2048)
dt = np.dtype(np.uint16)
u = torch.randint(0, 65536, shape, device="cuda", dtype=torch.int32).to(torch.uint16)
uo = torch.empty_like(u)
ps = [ndimage._Pass("uniform", ax, "reflect", size=size) for ax in (1, 2)]
for _ in range(4):
    ndimage._launch_correlate1d_pair(u, uo, shape, dt, ps[0], ps[1], 0.0)
torch.cuda.synchronize()
print(_lib.last_kernel())
