#!/usr/bin/env python
"""Benchmark of the ndfilters hot path (BASELINE.json metric: gaussian_filter
GVoxels/s on a 2048^3 fp32 volume at 1/2/4/8 B200, % of HBM roofline).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # reference CPU path
    torchrun --nproc-per-node N bench.py --gpus N ...        # N > 1 (one rank per GPU)

One "step" = one whole ``gaussian_filter`` over the synthetic volume (every 1-D
pass of scipy's chain; for N > 1 including the NCCL halo exchange).  ``value``
is device-resident throughput (inputs already in HBM), ``e2e`` is the same
public call fed from / drained to pinned HOST buffers inside the timed region.
The volume is the same for every N ("strong" scaling: fixed total work).
Prints exactly one JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "gaussian_filter_gvoxels_per_s"
UNIT = "GVoxels/s"

WORKLOADS = {
    # BASELINE.json `metric` (SURVEY.md 8(d) headline): fits one B200 (96 GiB)
    "headline": dict(shape=(2048, 2048, 2048), sigma=(2.0, 2.0, 2.0), mode="reflect",
                     desc="gaussian_filter sigma=2 mode=reflect on 2048^3 float32"),
    "headline_s4": dict(shape=(2048, 2048, 2048), sigma=(4.0, 4.0, 4.0), mode="reflect",
                        desc="gaussian_filter sigma=4 mode=reflect on 2048^3 float32"),
    # BASELINE.json configs[1]
    "c2": dict(shape=(1024, 1024, 1024), sigma=(4.0, 4.0, 4.0), mode="reflect",
               desc="gaussian_filter sigma=(4,4,4) mode=reflect on 1024^3 float32"),
    "small": dict(shape=(256, 256, 256), sigma=(2.0, 2.0, 2.0), mode="reflect",
                  desc="gaussian_filter sigma=2 mode=reflect on 256^3 float32 (debug)"),
}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def measured_traffic(kernel, voxels_per_launch):
    """DRAM bytes per launch of ``kernel`` from the committed ncu --set full
    capture (profiles/traffic.json: bytes per voxel), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            per_voxel = json.load(f)["dram_bytes_per_voxel"][kernel]
        return float(per_voxel) * voxels_per_launch
    except Exception:
        return None


# ---------------------------------------------------------------------------
# clocks sampling during the timed region
# ---------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
            time.sleep(0.3)          # let the first samples land before the timed region
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.1)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        try:
            rows = [r.strip().split(",") for r in open(self.path) if r.strip()]
            os.unlink(self.path)
        except Exception:
            return out
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1]))
                smax.append(float(r[2]))
                for nm, v in zip(names, r[5:9]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(smax)),
                       reasons=sorted(reasons), samples=len(sm))
        return out


# ---------------------------------------------------------------------------
# CPU reference arm: the reference's own path = dask threads x scipy.ndimage
# per halo'd chunk (dask is not installable here, so map_overlap is emulated
# exactly: halo = ceil(sigma*truncate), boundary "none", trim; verified
# bit-identical to whole-array scipy in tests/test_oracle.py)
# ---------------------------------------------------------------------------
def cpu_reference_run(wl, steps, warmup, budget_seconds=60.0):
    import scipy.ndimage as ndi
    from oracle import oracle as orc

    cores = len(os.sched_getaffinity(0))
    sigma, mode = wl["sigma"], wl["mode"]
    # bounded sample of the same workload: a cube sized so that warmup + steps
    # passes of it cost ~budget_seconds at ~0.012 GVox/s/core
    rate = 0.012e9 * cores * (2.0 / max(sigma)) ** 0.5
    side = int((rate * budget_seconds / max(1, steps + warmup)) ** (1.0 / 3) / 64.0) * 64
    side = int(min(max(side, 128), min(wl["shape"])))
    chunk = max(64, side // 4) if cores > 1 else side
    rng = np.random.default_rng(1234)
    a = rng.random((side, side, side), dtype=np.float32)
    depth = orc.gaussian_depth(sigma, 4.0, 3)

    def step():
        return orc.map_overlap_emulated(ndi.gaussian_filter, a, (chunk,) * 3, depth, "none",
                                        n_threads=cores, sigma=sigma, mode=mode)

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / max(1, steps)
    gvox = a.size / dt / 1e9
    sample = ("%d^3 float32 sub-volume of the workload, %d^3 chunks + halo %s, scipy.ndimage.gaussian_filter "
              "per chunk on a %d-thread pool (dask map_overlap emulated; dask not installable)"
              % (side, chunk, depth[0], cores))
    return gvox, dt, cores, sample


def run_reference(args, wl, rank, world):
    if rank != 0:
        return
    gvox, dt, cores, sample = cpu_reference_run(wl, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": gvox, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": wl["desc"], "sample": sample},
        "cpu_baseline": {"value": gvox, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": gvox, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------
# CUDA arm
# ---------------------------------------------------------------------------
def run_b200(args, wl, rank, world, local_rank):
    import torch

    import dask_image_b200 as b200
    from dask_image_b200 import _lib, ndimage
    from dask_image_b200.sharded import ShardedVolume, slab_bounds

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        # keep stdout for the one JSON line: NCCL's banner / debug lines go to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)
    shape = tuple(wl["shape"])
    sigma, mode = wl["sigma"], wl["mode"]
    voxels = int(np.prod(shape))
    warmup = max(args.warmup, 3)

    def barrier():
        if dist is not None:
            dist.barrier()

    # ---- device-resident: the public call on a (sharded) resident volume ---
    halo = max(int(4.0 * s + 0.5) for s in sigma)
    vol = ShardedVolume.random(shape, np.float32, seed=1234, device=dev, halo=halo)
    res = None

    def step():
        nonlocal res
        res = None                      # release the previous result first
        res = b200.ndfilters.gaussian_filter(vol, sigma, mode=mode)

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    _lib.launch_count(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    barrier()
    launches = _lib.launch_count()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    if dist is not None:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        lt = torch.tensor([launches], device=dev, dtype=torch.int64)
        dist.all_reduce(lt)
        launches = int(lt.item())
    ms_per_step = ms / args.steps
    value = voxels / (ms_per_step * 1e-3) / 1e9

    # ---- per-kernel launch durations (CUDA events, same stream, same slab) --
    n_loc = vol.n_local
    lshape = (n_loc,) + shape[1:]
    lvox = int(np.prod(lshape))
    x_t, out_t = vol.local, res.local
    passes = ndimage._gaussian_passes(3, sigma, 0, mode, 4.0)
    per_pass = []
    for ps in passes:
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
        ndimage._launch_correlate1d(x_t, out_t, lshape, np.dtype(np.float32), ps.axis, ps.weights, ps.mode, 0.0, 0)
        torch.cuda.synchronize()
        evs[0].record()
        for k in range(args.steps):
            ndimage._launch_correlate1d(x_t, out_t, lshape, np.dtype(np.float32), ps.axis, ps.weights, ps.mode,
                                        0.0, 0)
            evs[k + 1].record()
        torch.cuda.synchronize()
        t = float(np.mean([evs[k].elapsed_time(evs[k + 1]) for k in range(args.steps)]))
        per_pass.append((_lib.last_kernel(), ps.axis, len(ps.weights), t))
    peak, peak_src = measured_peaks()
    alg_bytes = 2 * 4 * lvox               # one read + one write of the f32 slab per pass
    by_kernel, passes_info = {}, []
    for name, ax, taps, t in per_pass:
        by_kernel.setdefault(name, []).append(t)
        passes_info.append({"kernel": name, "axis": ax, "taps": taps, "ms": t,
                            "gb_s": alg_bytes / (t * 1e-3) / 1e9})
    dom = max(by_kernel, key=lambda k: sum(by_kernel[k]))
    t_dom = float(np.mean(by_kernel[dom]))
    achieved = alg_bytes / (t_dom * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "frac_of_8000": achieved / 8000.0,
                "traffic": measured_traffic(dom, lvox), "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": t_dom,
                "launches_per_step": {k: len(v) for k, v in by_kernel.items()},
                "kernel_share_of_step": sum(by_kernel[dom]) / sum(t for _, _, _, t in per_pass)}

    # ---- end to end: pinned host in -> device -> public call -> pinned host out
    e2e = None
    if not args.no_e2e:
        try:
            s0, s1 = slab_bounds(shape[0], world)[rank]
            h_in = torch.empty(lshape, dtype=torch.float32).pin_memory()
            h_out = torch.empty(lshape, dtype=torch.float32).pin_memory()
            h_in.copy_(vol.local)                     # realistic content
            res = None
            vol = None
            x_t = out_t = None
            torch.cuda.empty_cache()

            # host volume (this rank's slab) streamed through the GPU: H2D | filter | D2H
            # overlapped on three streams, neighbour faces over NCCL for N > 1
            hv = b200.HostVolume(h_in, out=h_out, device=dev,
                                 global_shape=shape if world > 1 else None)

            def e2e_step():
                b200.ndfilters.gaussian_filter(hv, sigma, mode=mode)

            how = ("pinned host volume (one axis-0 slab per rank) -> ndfilters.gaussian_filter(HostVolume) -> "
                   "pinned host result: chunked H2D | filter | D2H pipeline on three streams%s; wall clock "
                   "with device sync, max over ranks" % (", NCCL halo faces" if world > 1 else ""))

            e2e_step()
            torch.cuda.synchronize()
            barrier()
            n_e2e = max(1, min(args.steps, 5))
            t0 = time.perf_counter()
            for _ in range(n_e2e):
                e2e_step()
            torch.cuda.synchronize()
            barrier()
            t_e2e = (time.perf_counter() - t0) / n_e2e
            if dist is not None:
                t = torch.tensor([t_e2e], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                t_e2e = float(t.item())
            e2e = {"value": voxels / t_e2e / 1e9, "unit": UNIT, "h2d_bytes_per_step": voxels * 4,
                   "d2h_bytes_per_step": voxels * 4, "ms_per_step": t_e2e * 1e3, "steps": n_e2e,
                   "how": how}
        except Exception as exc:  # e.g. not enough pinnable host memory
            e2e = {"value": None, "unit": UNIT, "error": repr(exc)[:200]}

    if rank != 0:
        return
    cpu = None
    if world == 1 and not args.no_cpu:
        gv, dt, cores, sample = cpu_reference_run(wl, 1, 1, budget_seconds=25.0)
        cpu = {"value": gv, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["desc"], "shape": list(shape), "sigma": list(sigma), "mode": mode,
                   "passes": len(passes),
                   "l2": "per-GPU slab (%.1f GiB) larger than L2; no flush needed" % (lvox * 4 / 2 ** 30),
                   "decomposition": ("axis-0 slabs (%d planes/GPU), NCCL send/recv halo exchange overlapped "
                                     "with interior compute" % n_loc) if world > 1 else "single GPU"},
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
        "clocks": clocks, "passes": passes_info,
        "op_level_roofline_frac": (8.0 * len(passes) * voxels / world / (ms_per_step * 1e-3) / 1e9) / peak,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer end-to-end leg")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    wl = WORKLOADS[args.workload or os.environ.get("B200_BENCH_WORKLOAD", "headline")]
    if args.impl == "reference":
        run_reference(args, wl, rank, world)
    else:
        run_b200(args, wl, rank, world, local_rank)
    if world > 1:
        import torch.distributed as dist

        if dist.is_initialized():
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
