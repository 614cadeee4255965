"""Step-by-step check of the IPC peer-halo path (run under torchrun, 2 ranks)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dask_image_b200 as b200  # noqa: E402
from dask_image_b200 import _lib, ndimage as nd, sharded  # noqa: E402

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")


def say(*a):
    print("[rank %d]" % rank, *a, flush=True)


# (0) single GPU, big geometry, local "peers"
if rank == 0:
    n, inner_shape = 512, (1024, 2048)
    x = torch.rand((n + 16,) + inner_shape, device=dev)
    w = nd._gaussian_weights(2.0, 0, 4.0)
    whole = torch.empty_like(x)
    nd._launch_correlate1d(x, whole, tuple(x.shape), np.dtype("f4"), 0, w, "reflect", 0.0, 0)
    out = torch.empty((n,) + inner_shape, device=dev)
    nd._launch_correlate1d_halo(x[8:8 + n], out, (n,) + inner_shape, np.dtype("f4"), w, "reflect", 0.0, 0,
                                x[0:8], x[8 + n:16 + n])
    torch.cuda.synchronize()
    say("step0 local big geometry:", _lib.last_kernel(), "equal:", bool(torch.equal(out, whole[8:8 + n])))
    del x, whole, out
    torch.cuda.empty_cache()
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
    shape = (256, 64, 256)
    vol = sharded.ShardedVolume.random(shape, np.float32, device=dev, halo=8)
    views = vol._peer_views(vol._buf)
    say("step1 views:", None if views is None else {k: (tuple(v[0].shape), str(v[0].device), v[1], v[2]) for k, v in views.items()})
    dist.barrier()
    name = "next" if rank == 0 else "prev"
    full, cap, n_p = views[name]
    # (2) kernel-level read of peer memory through the combine kernel
    mine = torch.zeros((4,) + shape[1:], device=dev)
    nd._launch_combine(mine, full[cap:cap + 4], np.dtype("f4"), _lib.EPI_STORE)
    torch.cuda.synchronize()
    ref = full[cap:cap + 4].to(dev)
    say("step2 kernel read of peer planes equal:", bool(torch.equal(mine, ref)))
    # (3) the sharded op through the peer path, small
    os.environ["B200_PEER_HALO"] = "1"
    got = b200.ndfilters.gaussian_filter(vol, 2.0)
    torch.cuda.synchronize()
    say("step3 small sharded gaussian done, kernel:", _lib.last_kernel())
    g = got.gather()
    one = b200.ndfilters.gaussian_filter(b200.B200Array(torch.from_numpy(vol.gather()).to(dev)), 2.0).to_numpy()
    say("step3 equal to 1 GPU:", bool(np.array_equal(g, one)))
    dist.barrier()
    dist.destroy_process_group()
