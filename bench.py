#!/usr/bin/env python
"""Benchmark of the ndfilters hot path (BASELINE.json metric: gaussian_filter
GVoxels/s on fp32 volumes, % of HBM roofline).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # reference CPU path

One "step" = one whole gaussian_filter over the synthetic volume (every 1-D
pass of scipy's chain).  ``value`` is device-resident throughput, ``e2e`` is
the same call fed from / drained to pinned HOST buffers inside the timed
region.  Prints exactly one JSON line on rank 0.
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

# name -> (per-GPU-or-global shape, sigma, mode, description)
WORKLOADS = {
    # BASELINE.json configs[1]
    "c2": dict(shape=(1024, 1024, 1024), sigma=(4.0, 4.0, 4.0), mode="reflect",
               desc="gaussian_filter sigma=(4,4,4) mode=reflect on 1024^3 float32"),
    # SURVEY.md 8(d) headline
    "headline": dict(shape=(2048, 2048, 2048), sigma=(2.0, 2.0, 2.0), mode="reflect",
                     desc="gaussian_filter sigma=2 mode=reflect on 2048^3 float32"),
    "headline_s4": dict(shape=(2048, 2048, 2048), sigma=(4.0, 4.0, 4.0), mode="reflect",
                        desc="gaussian_filter sigma=4 mode=reflect on 2048^3 float32"),
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
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.15)
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
def cpu_reference_run(wl, steps, warmup, target_seconds=12.0):
    import scipy.ndimage as ndi
    from oracle import oracle as orc

    cores = len(os.sched_getaffinity(0))
    sigma, mode = wl["sigma"], wl["mode"]
    # bounded sample of the same workload: a cube sized for ~target_seconds of
    # CPU work per step at ~0.012 GVox/s/core (SURVEY.md section 6 probes)
    rate = 0.012e9 * cores * (2.0 / max(sigma)) ** 0.5
    side = int(round((rate * target_seconds / max(1, steps + warmup)) ** (1.0 / 3) / 64.0)) * 64
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

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    shape = tuple(wl["shape"])
    sigma, mode = wl["sigma"], wl["mode"]
    voxels = int(np.prod(shape))

    if world > 1:
        from dask_image_b200.sharded import ShardedVolume

        vol = ShardedVolume.random(shape, np.float32, seed=1234)
        def step():
            return b200.ndfilters.gaussian_filter(vol, sigma, mode=mode)
        barrier = dist.barrier
    else:
        gen = torch.Generator(device=dev)
        gen.manual_seed(1234)
        x = b200.B200Array(torch.rand(shape, dtype=torch.float32, device=dev, generator=gen))
        out = x.empty_like()
        scratch = []

        def step():
            passes = ndimage._gaussian_passes(3, sigma, 0, mode, 4.0)
            ndimage._run_chain(x.tensor, x.shape, x.dtype, passes, 0.0, out.tensor, scratch=scratch)
            return out

        def barrier():
            return None

    for _ in range(max(args.warmup, 3)):
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
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    value = voxels / (ms_per_step * 1e-3) / 1e9

    # ---- per-kernel timing for the roofline (single GPU, same buffers) ------
    roofline = None
    passes_info = []
    if world == 1:
        passes = ndimage._gaussian_passes(3, sigma, 0, mode, 4.0)
        per_pass = []
        for pi, ps in enumerate(passes):
            src, dst = x.tensor, out.tensor      # timing only: contents do not matter
            evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
            ndimage._launch_correlate1d(src, dst, x.shape, x.dtype, ps.axis, ps.weights, ps.mode, 0.0, 0)
            torch.cuda.synchronize()
            evs[0].record()
            for k in range(args.steps):
                ndimage._launch_correlate1d(src, dst, x.shape, x.dtype, ps.axis, ps.weights, ps.mode, 0.0, 0)
                evs[k + 1].record()
            torch.cuda.synchronize()
            t = np.mean([evs[k].elapsed_time(evs[k + 1]) for k in range(args.steps)])
            per_pass.append((_lib.last_kernel(), ps.axis, len(ps.weights), float(t)))
        peak, peak_src = measured_peaks()
        alg_bytes = 2 * 4 * voxels           # one read + one write of the f32 volume per pass
        by_kernel = {}
        for name, ax, taps, t in per_pass:
            by_kernel.setdefault(name, []).append(t)
            passes_info.append({"kernel": name, "axis": ax, "taps": taps, "ms": t,
                                "gb_s": alg_bytes / (t * 1e-3) / 1e9})
        dom = max(by_kernel, key=lambda k: sum(by_kernel[k]))
        t_dom = float(np.mean(by_kernel[dom]))
        achieved = alg_bytes / (t_dom * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "frac_of_8000": achieved / 8000.0, "traffic": None,
                    "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes,
                    "launch_ms": t_dom}

    # ---- end to end: pinned host in -> device -> filter -> pinned host out ---
    e2e = None
    if world == 1 and not args.no_e2e:
        try:
            h_in = torch.empty(shape, dtype=torch.float32).pin_memory()
            h_out = torch.empty(shape, dtype=torch.float32).pin_memory()
            h_in.copy_(x.tensor)          # realistic content
            torch.cuda.synchronize()

            def e2e_step():
                d_in = b200.B200Array(h_in.to(dev, non_blocking=True))
                res = b200.ndfilters.gaussian_filter(d_in, sigma, mode=mode)
                h_out.copy_(res.tensor, non_blocking=True)

            # free the device-resident buffers of the first measurement
            e2e_step()
            torch.cuda.synchronize()
            n_e2e = max(1, min(args.steps, 5))
            t0 = time.perf_counter()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            for _ in range(n_e2e):
                e2e_step()
            a1.record()
            torch.cuda.synchronize()
            wall = (time.perf_counter() - t0) / n_e2e
            dev_ms = a0.elapsed_time(a1) / n_e2e
            t_e2e = max(wall, dev_ms * 1e-3)
            e2e = {"value": voxels / t_e2e / 1e9, "unit": UNIT, "h2d_bytes_per_step": voxels * 4,
                   "d2h_bytes_per_step": voxels * 4, "ms_per_step": t_e2e * 1e3, "steps": n_e2e}
        except Exception as exc:  # e.g. not enough pinnable host memory
            e2e = {"value": None, "unit": UNIT, "error": repr(exc)[:200]}

    if rank != 0:
        return
    cpu = None
    if world == 1 and not args.no_cpu:
        gv, dt, cores, sample = cpu_reference_run(wl, 1, 1)
        cpu = {"value": gv, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["desc"], "shape": list(shape), "sigma": list(sigma), "mode": mode,
                   "passes": 3, "l2": "inputs (%.1f GiB) larger than L2; no flush needed" % (voxels * 4 / 2 ** 30),
                   "decomposition": "axis-0 slabs, NCCL halo exchange" if world > 1 else "single GPU"},
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
        "clocks": clocks, "passes": passes_info,
        "op_level_roofline_frac": (24.0 * voxels / (ms_per_step * 1e-3) / 1e9) / measured_peaks()[0],
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
    wl = WORKLOADS[args.workload or os.environ.get("B200_BENCH_WORKLOAD", "c2")]
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
