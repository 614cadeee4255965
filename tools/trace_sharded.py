"""Phase timeline of one sharded gaussian_filter step (run under torchrun).
CUDA events on the compute stream between the engine's phases; printed per rank."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dask_image_b200 as b200  # noqa: E402
from dask_image_b200 import sharded  # noqa: E402

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
sigma = float(sys.argv[2]) if len(sys.argv) > 2 else 2.0
vol = sharded.ShardedVolume.random((n, n, n), np.float32, device=dev, halo=16)

marks = []
orig_axis0 = sharded.ShardedVolume._axis0_call
orig_exchange = sharded.ShardedVolume._exchange


def mark(name):
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    marks.append((name, e))


def traced_exchange(self, buf, lo, hi, ring):
    mark("pre-exchange")
    r = orig_exchange(self, buf, lo, hi, ring)
    mark("exchange-posted")
    return r


def traced_axis0(self, src, lo, hi, ring, call, cache=None):
    def call2(p0, p1, pad):
        call(p0, p1, pad)
        mark("axis0[%d:%d]" % (p0, p1))
    orig_axis0(self, src, lo, hi, ring, call2, cache)


sharded.ShardedVolume._exchange = traced_exchange
sharded.ShardedVolume._axis0_call = traced_axis0
for it in range(6):
    marks.clear()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    mark("start")
    res = b200.ndfilters.gaussian_filter(vol, sigma)
    mark("end")
    torch.cuda.synchronize()
    del res
if world > 1:
    dist.barrier()
for r in range(world):
    if r == rank:
        t0 = marks[0][1]
        print("rank %d: " % rank + "  ".join("%s@%.3f" % (nm, t0.elapsed_time(e)) for nm, e in marks), flush=True)
    if world > 1:
        dist.barrier()
if world > 1:
    dist.destroy_process_group()
