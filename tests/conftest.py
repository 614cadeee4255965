import os
import sys

import pytest

# the slab-engine tests cut SMALL volumes on purpose: switch the replicate-small-volumes rule off
# (must be set before dask_image_b200 is imported; spawned NCCL / gloo workers inherit it)
os.environ.setdefault("B200_REPLICATE_BELOW", "0")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device visible")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def pytest_sessionfinish(session, exitstatus):
    """Achieved maximum errors (relative to the output / data scale) of every float parity check of a GPU
    session -> gpurun_out/parity_errors.json (copied to profiles/ for the record)."""
    try:
        import json

        from _helpers import ACHIEVED

        import torch

        if not ACHIEVED or not torch.cuda.is_available():
            return
        out = os.path.join(ROOT, "gpurun_out")
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity_errors.json"), "w") as f:
            json.dump({"max_abs_err_over_scale": dict(sorted(ACHIEVED.items())),
                       "bound": "1e-5 |ref| + 2e-6 scale (float32), 1e-12 |ref| + 2e-13 scale (float64)"}, f, indent=1)
    except Exception:
        pass
