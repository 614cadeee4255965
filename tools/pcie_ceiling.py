"""Host <-> device copy ceiling of the box: the roofline of the end-to-end (host buffers in, host
buffers out) metric.  One process per GPU (run under torchrun for N > 1); every rank copies a pinned
host buffer to its GPU and another device buffer back to pinned host memory, one cudaMemcpyAsync per
256 MiB chunk, on one stream per direction, both directions at once -- exactly the traffic pattern of
HostVolume -- and rank 0 prints the aggregate GB/s per direction (max time over ranks).

    python tools/pcie_ceiling.py [--gib 4] [--bind 0|1]
    torchrun --nproc-per-node N tools/pcie_ceiling.py
"""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dask_image_b200 import _numa  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gib", type=float, default=4.0, help="bytes per direction per rank and repetition")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--chunk-mib", type=int, default=256)
    ap.add_argument("--bind", type=int, default=1)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    numa = _numa.bind_host_to_device(dev) if args.bind else _numa.device_numa_info(dev)
    numa.pop("local_cpus", None)
    n = int(args.gib * (1 << 30))
    chunk = args.chunk_mib << 20
    h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    h_out = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    h_in.fill_(1)
    h_out.fill_(0)
    d_in = torch.empty(n, dtype=torch.uint8, device=dev)
    d_out = torch.ones(n, dtype=torch.uint8, device=dev)
    s_up, s_down = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def run(up, down):
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(args.reps):
            for o in range(0, n, chunk):
                if up:
                    with torch.cuda.stream(s_up):
                        d_in[o:o + chunk].copy_(h_in[o:o + chunk], non_blocking=True)
                if down:
                    with torch.cuda.stream(s_down):
                        h_out[o:o + chunk].copy_(d_out[o:o + chunk], non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        return world * n * args.reps / dt / 1e9   # aggregate GB/s per direction

    run(True, True)   # warm-up
    res = {"n_gpus": world, "gib_per_direction_per_rank": args.gib, "chunk_mib": args.chunk_mib,
           "bind": bool(args.bind), "h2d_only_gbs": run(True, False), "d2h_only_gbs": run(False, True),
           "both_gbs_per_direction": run(True, True)}
    infos = [numa]
    if dist is not None:
        infos = [None] * world
        dist.all_gather_object(infos, numa)
    if rank == 0:
        res["ranks"] = infos
        print(json.dumps(res), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
