"""Device-resident timing of every BASELINE.json config on the visible GPUs
(run under torchrun for N > 1).  One JSON line per config on rank 0."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dask_image_b200 as b200  # noqa: E402
from dask_image_b200 import _lib  # noqa: E402
from dask_image_b200.sharded import ShardedVolume  # noqa: E402

PEAK = 6545.9


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    only = set(sys.argv[1:])
    scale = int(os.environ.get("B200_CFG_SCALE", "1"))       # divide every edge (debug)
    rng = np.random.default_rng(1234)
    w9 = rng.random((9, 9, 9), dtype=np.float32)
    f = b200.ndfilters
    cfgs = [
        ("C1 gaussian s=2 4096^2 f32", (4096 // scale, 4096 // scale), np.float32, 16,
         lambda v: f.gaussian_filter(v, 2.0)),
        ("C2 gaussian s=4 1024^3 f32", (1024 // scale,) * 3, np.float32, 24,
         lambda v: f.gaussian_filter(v, (4.0, 4.0, 4.0))),
        ("C3 uniform 7 2048^3 u16", (2048 // scale,) * 3, np.uint16, 12,
         lambda v: f.uniform_filter(v, 7)),
        ("C4a ggm s=3 2048^3 f32", (2048 // scale,) * 3, np.float32, 72,
         lambda v: f.gaussian_gradient_magnitude(v, 3.0)),
        ("C4b glaplace s=3 2048^3 f32", (2048 // scale,) * 3, np.float32, 72,
         lambda v: f.gaussian_laplace(v, 3.0)),
        ("C5 convolve 9^3 wrap 1536^3 f32", (1536 // scale,) * 3, np.float32, 8,
         lambda v: f.convolve(v, w9, mode="wrap")),
        ("M maximum 7 2048^3 u16", (2048 // scale,) * 3, np.uint16, 12,
         lambda v: f.maximum_filter(v, size=7)),
        ("H gaussian s=2 2048^3 f32", (2048 // scale,) * 3, np.float32, 24,
         lambda v: f.gaussian_filter(v, 2.0)),
        ("H4 gaussian s=4 2048^3 f32", (2048 // scale,) * 3, np.float32, 24,
         lambda v: f.gaussian_filter(v, 4.0)),
        # integer images: the exact tiled integer kernels (instruction-issue bound, not HBM bound)
        ("I16 gaussian s=2 2048^3 u16", (2048 // scale,) * 3, np.uint16, 12,
         lambda v: f.gaussian_filter(v, 2.0)),
        ("I8 gaussian s=2 2048^3 u8", (2048 // scale,) * 3, np.uint8, 6,
         lambda v: f.gaussian_filter(v, 2.0)),
        ("I16g ggm s=2 1024^3 u16", (1024 // scale,) * 3, np.uint16, 36,
         lambda v: f.gaussian_gradient_magnitude(v, 2.0)),
    ]
    for name, shape, dt, bpv, fn in cfgs:
        if only and name.split()[0] not in only:
            continue
        try:
            vol = ShardedVolume.random(shape, dt, seed=1234, device=dev, halo=16)
            res = fn(vol)
            del res
            torch.cuda.synchronize()
            reps = 1 if name.startswith("C5") else 3
            _lib.launch_count(reset=True)
            if dist:
                dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                res = fn(vol)
                del res
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            if dist:
                t = torch.tensor([ms], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t.item())
            vox = float(np.prod(shape))
            if rank == 0:
                print(json.dumps({
                    "config": name, "n_gpus": world, "ms": round(ms, 3),
                    "gvox_s": round(vox / ms / 1e6, 2),
                    "alg_GBs_per_gpu": round(vox * bpv / world / ms / 1e6, 1),
                    "frac_of_measured_peak": round(vox * bpv / world / ms / 1e6 / PEAK, 3),
                    "launches_rank0": _lib.launch_count() // reps, "last_kernel": _lib.last_kernel()}),
                    flush=True)
            del vol
            torch.cuda.empty_cache()
        except Exception as exc:
            if rank == 0:
                print(json.dumps({"config": name, "error": repr(exc)[:300]}), flush=True)
            torch.cuda.empty_cache()
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
