"""Smallest use of the fused integer box pair (for compute-sanitizer): python tools/box_min.py [size=7]"""
import os
import sys

import numpy as np
import scipy.ndimage as ndi

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dask_image_b200 as b200  # noqa: E402
from dask_image_b200 import _lib  # noqa: E402

size = int(sys.argv[1]) if len(sys.argv) > 1 else 7
for dtype, shape in ((np.uint16, (2, 40, 64)), (np.uint16, (3, 70, 520)), (np.uint8, (2, 40, 64))):
    a = np.random.default_rng(1).integers(0, np.iinfo(dtype).max + 1, shape).astype(dtype)
    got = b200.ndimage.uniform_filter(b200.B200Array.from_numpy(a), size).to_numpy()
    ref = ndi.uniform_filter(a, size)
    print(np.dtype(dtype).name, shape, _lib.last_kernel(), "equal" if np.array_equal(got, ref) else
          "DIFFERENT: %d voxels" % int((got != ref).sum()), flush=True)
