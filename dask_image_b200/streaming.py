"""``HostVolume`` -- filter a volume that lives in HOST memory, streamed through
one GPU.

The reference's users hold host (numpy-backed) dask arrays; per chunk dask
would copy a halo'd block to the device, filter it and copy the result back
(dask_image/ndfilters/_gaussian.py:70-81 with a device chunk type).  This
module is the B200-native form of that loop for a whole volume: the input is
uploaded ONCE, in axis-0 chunks, and three streams overlap

    H2D of chunk k+1  |  filter of chunk k  |  D2H of chunk k-1

so the end-to-end time is bounded by the PCIe link (both directions busy at
once) instead of upload + compute + download in sequence.  No halo is ever
re-uploaded and nothing is trimmed: a chunk is filtered in place inside the
resident input through the slab contract of the C ABI (``pad_lo`` / ``pad_hi``:
the planes around it are simply resident), so the result is bit-identical to
filtering the whole resident volume.

When the volume fits the device (the benchmark case) the input stays resident
while it is filtered.  A single-process volume that does NOT fit (or is given a
smaller ``max_device_bytes``) is filtered out of core: axis-0 blocks travel with
their halo planes through two device windows (``_run_out_of_core``).  A slab of
a SHARDED volume must fit its GPU (~160 GB per rank); a larger one raises
``MemoryError`` up front -- spread the volume over more ranks.

Supported: every plan whose axis-0 passes read the plan's INPUT (gaussian
filters and their derivative compositions, uniform_filter, laplace, N-D
correlate / convolve).  Plans that filter along axis 0 after another pass
(sobel / prewitt along a different axis) run un-pipelined on the whole volume.
"""
import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from . import ndimage as _nd
from ._array import _TORCH_TO_NP, torch_dtype
from .sharded import ShardedVolume, slab_bounds

__all__ = ["HostVolume", "FUNCTIONS"]


def _as_host_tensor(a):
    if isinstance(a, torch.Tensor):
        if a.device.type != "cpu":
            raise TypeError("HostVolume wraps host memory; use B200Array for device tensors")
        return a if a.is_contiguous() else a.contiguous()
    a = np.ascontiguousarray(a)
    torch_dtype(a.dtype)
    if a.dtype in (np.uint16, np.uint32, np.uint64):
        signed = {np.dtype(np.uint16): np.int16, np.dtype(np.uint32): np.int32,
                  np.dtype(np.uint64): np.int64}[a.dtype]
        return torch.from_numpy(a.view(signed)).view(torch_dtype(a.dtype))
    return torch.from_numpy(a)


class HostVolume:
    """A C-contiguous N-D volume in host memory (pinned memory gives full PCIe
    speed and true overlap; pageable memory works, slower).

    Parameters
    ----------
    array : numpy.ndarray or CPU torch.Tensor
        The whole volume -- or, with ``global_shape``, this rank's axis-0 slab
        of a volume spread over the ranks of ``group`` (one process per GPU;
        slabs as :func:`dask_image_b200.sharded.slab_bounds` cuts them).
    out : optional CPU tensor / array of the same shape and dtype that receives
        the result of the next filter (otherwise a pinned one is allocated)
    device : CUDA device to stream through (default: current)
    chunk_bytes : approximate size of one upload / compute / download unit (B200 / PCIe 5 scan,
        tools/e2e_scan.py: 128-512 MiB is the plateau)
    max_device_bytes : device memory the filter may use (default: what is free).  A volume (one process, not
        sharded) that does not fit is filtered OUT OF CORE: in axis-0 blocks that are uploaded with their halo
        planes, filtered and downloaded on three streams, two blocks in flight -- same result, bit for bit.
    """

    def __init__(self, array, out=None, device=None, chunk_bytes=256 << 20, global_shape=None, group=None,
                 max_device_bytes=None):
        self.tensor = _as_host_tensor(array)
        if self.tensor.dtype not in _TORCH_TO_NP:
            raise RuntimeError("data type not supported")
        self.out = None if out is None else _as_host_tensor(out)
        if self.out is not None and (self.out.shape != self.tensor.shape or self.out.dtype != self.tensor.dtype):
            raise RuntimeError("output shape / dtype not correct")
        if device is None:
            if not torch.cuda.is_available():
                raise RuntimeError("dask_image_b200 needs a CUDA device (sm_100a); none is visible "
                                   "and there is no CPU fallback")
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = torch.device(device)
        self.chunk_bytes = int(chunk_bytes)
        # device memory the filter may use; None = what is free on the device when the filter runs
        self.max_device_bytes = None if max_device_bytes is None else int(max_device_bytes)
        self.group = group
        self.world = self.rank = None
        self.global_shape = None
        if global_shape is not None:
            if not (dist.is_available() and dist.is_initialized()):
                raise RuntimeError("a sharded HostVolume needs torch.distributed to be initialised")
            self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
            self.global_shape = tuple(int(v) for v in global_shape)
            s0, s1 = slab_bounds(self.global_shape[0], self.world)[self.rank]
            if tuple(self.tensor.shape) != (s1 - s0,) + self.global_shape[1:]:
                raise ValueError("rank %d must hold planes [%d, %d) of %s" % (self.rank, s0, s1, self.global_shape))

    @classmethod
    def pinned_empty(cls, shape, dtype, **kw):
        return cls(torch.empty(tuple(shape), dtype=torch_dtype(dtype)).pin_memory(), **kw)

    @property
    def shape(self):
        """Global shape of the volume."""
        return self.global_shape if self.global_shape is not None else tuple(int(s) for s in self.tensor.shape)

    @property
    def ndim(self):
        return self.tensor.dim()

    @property
    def dtype(self):
        return _TORCH_TO_NP[self.tensor.dtype]

    @property
    def size(self):
        return int(np.prod(self.shape, dtype=np.int64))

    def to_numpy(self):
        """This process's planes as a numpy array (the whole volume unless sharded)."""
        t = self.tensor
        if t.dtype in (torch.uint16, torch.uint32, torch.uint64):
            return t.view(torch.uint8).numpy().view(self.dtype).reshape(tuple(t.shape))
        return t.numpy()

    def __array__(self, dtype=None, copy=None):
        a = self.to_numpy()
        return a if dtype is None else a.astype(dtype, copy=False)

    def __repr__(self):
        return "HostVolume(shape=%s, dtype=%s, pinned=%s, via=%s%s)" % (
            self.shape, self.dtype, self.tensor.is_pinned(), self.device,
            "" if self.world is None else ", rank %d/%d" % (self.rank, self.world))

    def _result(self, h_out):
        return HostVolume(h_out, device=self.device, chunk_bytes=self.chunk_bytes,
                          global_shape=self.global_shape, group=self.group, max_device_bytes=self.max_device_bytes)

    # ------------------------------------------------------------------
    def _run_plan(self, plan):
        h_in = self.tensor
        h_out = self.out if self.out is not None else torch.empty_like(h_in).pin_memory()
        result = self._result(h_out)
        if self.ndim == 0:
            raise RuntimeError("0-d arrays are not supported")
        if self.size == 0:
            return result
        dev, dt = self.device, self.dtype
        n0 = int(h_in.shape[0])                 # planes held by this process
        tail = tuple(int(v) for v in h_in.shape[1:])
        plane_bytes = max(1, int(np.prod(tail, dtype=np.int64)) * dt.itemsize)
        sharded = self.world is not None and self.world > 1

        # reach of the plan along axis 0; plans that filter along axis 0 after
        # another pass cannot be chunked without halos of intermediates
        lo = hi = 0
        streamable = True
        ring = False
        if plan.nd is not None:
            w, origins, mode = plan.nd
            lo = w.shape[0] // 2 + origins[0]
            hi = w.shape[0] - 1 - lo
            ring = mode in ("wrap", "grid-wrap")
        for passes, _ in plan.chains:
            for i, ps in enumerate(passes):
                if ps.axis == 0:
                    if i > 0:
                        streamable = False
                    L = len(ps.weights) if ps.kind == "corr" else ps.size
                    l0 = L // 2 + ps.origin
                    lo, hi = max(lo, l0), max(hi, L - 1 - l0)
                    ring = ring or ps.mode in ("wrap", "grid-wrap")
        if not plan.chains and plan.nd is None:
            h_out.copy_(h_in)
            return result
        # chunking decisions must agree on every rank: derive them from the thinnest slab
        nmin = n0
        if sharded:
            nmin = min(e - s for s, e in slab_bounds(self.global_shape[0], self.world))
        C = max(1, min(nmin, self.chunk_bytes // plane_bytes))
        C = max(C, lo, hi, 1)
        chunks = [(s, min(n0, s + C)) for s in range(0, n0, C)]
        pipelined = streamable and -(-nmin // C) >= 3

        # neighbours along axis 0 (other ranks; the ring closes for mode='wrap')
        has_prev = has_next = False
        prev = nxt = None
        if sharded:
            if max(lo, hi) > nmin:
                raise ValueError("halo depth %d exceeds the thinnest slab (%d planes on %d ranks)"
                                 % (max(lo, hi), nmin, self.world))
            has_prev = (ring or self.rank > 0) and lo > 0
            has_next = (ring or self.rank < self.world - 1) and hi > 0
            prev, nxt = (self.rank - 1) % self.world, (self.rank + 1) % self.world
            if self.group is not None:
                prev, nxt = dist.get_global_rank(self.group, prev), dist.get_global_rank(self.group, nxt)

        # does the slab fit the device?  (resident input + three output chunks + scratch chunks when pipelined;
        # input + output + scratch volumes otherwise)
        cap_lo, cap_hi = (lo if has_prev else 0), (hi if has_next else 0)
        need_bytes = ((cap_lo + n0 + cap_hi + 5 * C) if pipelined else 4 * n0) * plane_bytes
        free_bytes = (torch.cuda.mem_get_info(dev)[0] + torch.cuda.memory_reserved(dev)
                      - torch.cuda.memory_allocated(dev))
        if self.max_device_bytes is not None:
            free_bytes = min(free_bytes, self.max_device_bytes)
        if need_bytes > free_bytes:
            if not sharded and streamable:
                self._run_out_of_core(plan, h_in, h_out, lo, hi, ring, free_bytes)
                return result
            raise MemoryError("HostVolume: %.1f GiB of device memory needed, %.1f GiB available on %s -- %s"
                              % (need_bytes / 2 ** 30, free_bytes / 2 ** 30, dev,
                                 "spread the volume over more ranks (global_shape=...)" if sharded else
                                 "this filter runs along axis 0 after another pass and cannot be cut into blocks"))

        with torch.cuda.device(dev):
            cur = torch.cuda.current_stream(dev)
            if not pipelined:
                # un-pipelined: upload, filter the resident volume, download
                if sharded:
                    vol = ShardedVolume.from_local(h_in.to(dev, non_blocking=True), self.global_shape, self.group,
                                                   halo=max(lo, hi))
                    h_out.copy_(vol._run_plan(plan).local, non_blocking=True)
                else:
                    from ._array import B200Array

                    d_in = B200Array(h_in.to(dev, non_blocking=True))
                    d_out = _nd._execute(plan, d_in, d_in.empty_like())
                    h_out.copy_(d_out.tensor, non_blocking=True)
                cur.synchronize()
                return result

            d_full = torch.empty((cap_lo + n0 + cap_hi,) + tail, dtype=h_in.dtype, device=dev)
            d_in = d_full[cap_lo:cap_lo + n0]
            nslot = 3
            d_out = [torch.empty((C,) + tail, dtype=h_in.dtype, device=dev) for _ in range(nslot)]
            need = 0
            for passes, op in plan.chains:
                nsteps = len(_nd._chain_steps(passes, (C,) + tail, dt, axis0_local=False, cval=plan.cval))
                for tgt in _nd._chain_targets(nsteps, op):
                    if tgt != "out":
                        need = max(need, tgt + 1)
            scratch = [torch.empty((C,) + tail, dtype=h_in.dtype, device=dev) for _ in range(need)]
            s_up, s_cmp, s_down = (torch.cuda.Stream(dev) for _ in range(3))
            for s in (s_up, s_cmp, s_down):
                s.wait_stream(cur)
            ev_up = [torch.cuda.Event() for _ in chunks]
            ev_cmp = [torch.cuda.Event() for _ in chunks]
            ev_down = [None] * nslot
            halo_reqs = {"down": None, "up": None}

            def face(t):
                return t.view(torch.uint8)

            def post_exchange(which):
                """'down': my first planes -> prev's upper pad, next's first planes ->
                my upper pad (posted once chunk 0 is uploaded).  'up': my last
                planes -> next's lower pad, prev's last planes -> my lower pad
                (posted once the last chunk is uploaded).  Same order on every
                rank; the NCCL stream picks the uploads up from ``s_up``."""
                ops = []
                if which == "down":
                    if (ring or self.rank > 0) and hi > 0:
                        ops.append(dist.P2POp(dist.isend, face(d_in[0:hi]), prev, self.group, tag=0))
                    if has_next:
                        ops.append(dist.P2POp(dist.irecv, face(d_full[cap_lo + n0:cap_lo + n0 + hi]), nxt, self.group,
                                              tag=0))
                else:
                    if (ring or self.rank < self.world - 1) and lo > 0:
                        ops.append(dist.P2POp(dist.isend, face(d_in[n0 - lo:n0]), nxt, self.group, tag=1))
                    if has_prev:
                        ops.append(dist.P2POp(dist.irecv, face(d_full[0:lo]), prev, self.group, tag=1))
                if ops:
                    with torch.cuda.stream(s_up):
                        halo_reqs[which] = dist.batch_isend_irecv(ops)

            def compute(j, slot, uploaded_end, scr):
                c0, c1 = chunks[j]
                n = c1 - c0
                # planes resident around the chunk: the uploaded part of the slab,
                # the previous rank's face below chunk 0, the next rank's face above
                # the slab once all of it is up
                top = uploaded_end == n0
                pad = (cap_lo if j == 0 else c0, (uploaded_end - c1) + (cap_hi if top else 0))
                src = d_in[c0:c1]
                dst = d_out[slot][:n]
                shape = (n,) + tail
                with torch.cuda.stream(s_cmp):
                    s_cmp.wait_event(ev_up[min(len(chunks) - 1, _chunk_of(chunks, uploaded_end - 1))])
                    if j == 0 and halo_reqs["up"] is not None:
                        for rq in halo_reqs["up"]:
                            rq.wait()
                    if top and cap_hi and halo_reqs["down"] is not None:
                        for rq in halo_reqs["down"]:
                            rq.wait()
                    if ev_down[slot] is not None:
                        s_cmp.wait_event(ev_down[slot])
                    if plan.extreme is not None:
                        fp, origins, mode, is_max = plan.extreme
                        _nd._launch_minmax_nd(src, dst, shape, dt, fp, origins, is_max, mode, plan.cval, pad)
                    elif plan.nd is not None:
                        w, origins, mode = plan.nd
                        _nd._launch_correlate_nd(src, dst, shape, dt, w, origins, mode, plan.cval, pad)
                    else:
                        if plan.zero_out:
                            dst.zero_()
                        prov = {}
                        for passes, op in plan.chains:
                            if passes:
                                _nd._run_chain(src, shape, dt, passes, plan.cval, dst, final_epilogue=op,
                                               pad=pad, scratch=scr, prov=prov)
                            else:
                                _nd._launch_combine(dst, src, dt, op)
                    ev_cmp[j].record(s_cmp)
                with torch.cuda.stream(s_down):
                    s_down.wait_event(ev_cmp[j])
                    h_out[c0:c1].copy_(dst, non_blocking=True)
                    ev_down[slot] = torch.cuda.Event()
                    ev_down[slot].record(s_down)

            # chunk 0 reaches below the slab: the last planes of the volume for a
            # periodic one, the previous rank's last planes for a sharded one --
            # either way it is filtered at the end
            defer0 = ring or has_prev
            order = list(range(len(chunks)))
            if defer0:
                order = order[1:] + [0]
            pending = list(order)
            issued = 0
            for k, (c0, c1) in enumerate(chunks):
                with torch.cuda.stream(s_up):
                    d_in[c0:c1].copy_(h_in[c0:c1], non_blocking=True)
                    ev_up[k].record(s_up)
                uploaded_end = c1
                if sharded and k == 0:
                    post_exchange("down")
                if sharded and k == len(chunks) - 1:
                    post_exchange("up")
                while pending:
                    j = pending[0]
                    j0, j1 = chunks[j]
                    if j == 0 and defer0:
                        ready = uploaded_end == n0
                    elif j1 + hi <= n0:
                        ready = uploaded_end >= j1 + hi
                    else:
                        ready = uploaded_end == n0     # reaches the slab's upper face
                    if not ready:
                        break
                    pending.pop(0)
                    nj = j1 - j0
                    compute(j, issued % nslot, uploaded_end, scratch if nj == C else [t[:nj] for t in scratch])
                    issued += 1
            for s in (s_up, s_cmp, s_down):
                cur.wait_stream(s)
            cur.synchronize()
        return result


def _run_out_of_core(self, plan, h_in, h_out, lo, hi, ring, budget_bytes):
    """Filter a host volume that does not fit the device: axis-0 blocks of B planes, each uploaded with
    the ``lo`` / ``hi`` halo planes around it (the volume's own planes; the opposite end's for mode='wrap';
    none at a boundary of the volume, where the kernels fold by ``mode``), filtered through the slab
    contract of the C ABI (``pad_lo`` / ``pad_hi`` = the halo planes resident around the block) and
    downloaded -- upload of block k+1, filter of block k and download of block k-1 overlap on three
    streams, two windows and two result buffers alternate.  Every voxel sees exactly the operation sequence
    of the resident path: bit-identical result.  Extra PCIe traffic: (lo + hi) / B of the volume."""
    dev, dt = self.device, self.dtype
    n0 = int(h_in.shape[0])
    tail = tuple(int(v) for v in h_in.shape[1:])
    plane_bytes = max(1, int(np.prod(tail, dtype=np.int64)) * dt.itemsize)
    need = 0
    for passes, op in plan.chains:
        nsteps = len(_nd._chain_steps(passes, (2,) + tail, dt, axis0_local=False, cval=plan.cval))
        for tgt in _nd._chain_targets(nsteps, op):
            if tgt != "out":
                need = max(need, tgt + 1)
    planes = int(budget_bytes // plane_bytes)
    B = (planes - 2 * (lo + hi)) // (4 + need)
    if B < max(lo, hi, 1):
        raise MemoryError("HostVolume: %.2f GiB of device memory cannot hold two blocks of %d planes plus halos"
                          % (budget_bytes / 2 ** 30, max(lo, hi, 1)))
    B = min(B, n0)
    blocks = [(s, min(n0, s + B)) for s in range(0, n0, B)]
    with torch.cuda.device(dev):
        cur = torch.cuda.current_stream(dev)
        wins = [torch.empty((lo + B + hi,) + tail, dtype=h_in.dtype, device=dev) for _ in range(2)]
        outs = [torch.empty((B,) + tail, dtype=h_in.dtype, device=dev) for _ in range(2)]
        scratch = [torch.empty((B,) + tail, dtype=h_in.dtype, device=dev) for _ in range(need)]
        s_up, s_cmp, s_down = (torch.cuda.Stream(dev) for _ in range(3))
        for st in (s_up, s_cmp, s_down):
            st.wait_stream(cur)
        ev_cmp, ev_down = [None, None], [None, None]
        for k, (b0, b1) in enumerate(blocks):
            n, slot = b1 - b0, k % 2
            w = wins[slot]
            lo_eff = lo if (b0 > 0 or ring) else 0
            hi_eff = hi if ring else min(hi, n0 - b1)
            with torch.cuda.stream(s_up):
                if ev_cmp[slot] is not None:
                    s_up.wait_event(ev_cmp[slot])           # the filter of block k-2 has read this window
                w[lo:lo + n].copy_(h_in[b0:b1], non_blocking=True)
                if lo_eff:
                    if b0 >= lo:
                        w[0:lo].copy_(h_in[b0 - lo:b0], non_blocking=True)
                    else:                                     # first block of a periodic volume
                        for j in range(lo):
                            w[j].copy_(h_in[(b0 - lo + j) % n0], non_blocking=True)
                if hi_eff:
                    if b1 + hi_eff <= n0:
                        w[lo + n:lo + n + hi_eff].copy_(h_in[b1:b1 + hi_eff], non_blocking=True)
                    else:                                     # last block of a periodic volume
                        for j in range(hi_eff):
                            w[lo + n + j].copy_(h_in[(b1 + j) % n0], non_blocking=True)
                ev_up = torch.cuda.Event()
                ev_up.record(s_up)
            src, dst, shape, pad = w[lo:lo + n], outs[slot][:n], (n,) + tail, (lo_eff, hi_eff)
            with torch.cuda.stream(s_cmp):
                s_cmp.wait_event(ev_up)
                if ev_down[slot] is not None:
                    s_cmp.wait_event(ev_down[slot])          # the download of block k-2 has left this buffer
                if plan.extreme is not None:
                    fp, origins, mode, is_max = plan.extreme
                    _nd._launch_minmax_nd(src, dst, shape, dt, fp, origins, is_max, mode, plan.cval, pad)
                elif plan.nd is not None:
                    wts, origins, mode = plan.nd
                    _nd._launch_correlate_nd(src, dst, shape, dt, wts, origins, mode, plan.cval, pad)
                else:
                    if plan.zero_out:
                        dst.zero_()
                    prov = {}
                    scr = scratch if n == B else [t[:n] for t in scratch]
                    for passes, op in plan.chains:
                        if passes:
                            _nd._run_chain(src, shape, dt, passes, plan.cval, dst, final_epilogue=op, pad=pad,
                                           scratch=scr, prov=prov)
                        else:
                            _nd._launch_combine(dst, src, dt, op)
                ev_cmp[slot] = torch.cuda.Event()
                ev_cmp[slot].record(s_cmp)
            with torch.cuda.stream(s_down):
                s_down.wait_event(ev_cmp[slot])
                h_out[b0:b1].copy_(dst, non_blocking=True)
                ev_down[slot] = torch.cuda.Event()
                ev_down[slot].record(s_down)
        for st in (s_up, s_cmp, s_down):
            cur.wait_stream(st)
        cur.synchronize()


HostVolume._run_out_of_core = _run_out_of_core


def _chunk_of(chunks, plane):
    for k, (c0, c1) in enumerate(chunks):
        if c0 <= plane < c1:
            return k
    return len(chunks) - 1


def _check(vol):
    if not isinstance(vol, HostVolume):
        raise TypeError("expected a HostVolume")
    if vol.ndim == 0:
        raise RuntimeError("0-d arrays are not supported")
    return vol


def gaussian_filter(input, sigma, order=0, mode="reflect", cval=0.0, truncate=4.0, *, radius=None, axes=None):
    vol = _check(input)
    return vol._run_plan(_nd._cached_plan(_nd._plan_gaussian_filter, vol.ndim, (sigma, order, mode, cval, truncate, radius, axes,), {}))


def uniform_filter(input, size=3, mode="reflect", cval=0.0, origin=0, *, axes=None):
    vol = _check(input)
    return vol._run_plan(_nd._cached_plan(_nd._plan_uniform_filter, vol.ndim, (size, mode, cval, origin, axes,), {}))


def gaussian_gradient_magnitude(input, sigma, mode="reflect", cval=0.0, *, axes=None, **kwargs):
    vol = _check(input)
    return vol._run_plan(_nd._cached_plan(_nd._plan_gaussian_gradient_magnitude, vol.ndim, (sigma, mode, cval, axes,), kwargs))


def gaussian_laplace(input, sigma, mode="reflect", cval=0.0, *, axes=None, **kwargs):
    vol = _check(input)
    return vol._run_plan(_nd._cached_plan(_nd._plan_gaussian_laplace, vol.ndim, (sigma, mode, cval, axes,), kwargs))


def correlate(input, weights, mode="reflect", cval=0.0, origin=0, *, axes=None):
    vol = _check(input)
    return vol._run_plan(_nd._plan_correlate_nd(vol.ndim, vol.dtype, weights, mode, cval, origin, False, axes))


def convolve(input, weights, mode="reflect", cval=0.0, origin=0, *, axes=None):
    vol = _check(input)
    return vol._run_plan(_nd._plan_correlate_nd(vol.ndim, vol.dtype, weights, mode, cval, origin, True, axes))


def laplace(input, mode="reflect", cval=0.0, *, axes=None):
    vol = _check(input)
    return vol._run_plan(_nd._cached_plan(_nd._plan_laplace, vol.ndim, (mode, cval, axes,), {}))


def prewitt(input, axis=-1, mode="reflect", cval=0.0):
    vol = _check(input)
    return vol._run_plan(_nd._cached_plan(_nd._plan_prewitt, vol.ndim, (axis, mode, cval,), {}))


def sobel(input, axis=-1, mode="reflect", cval=0.0):
    vol = _check(input)
    return vol._run_plan(_nd._cached_plan(_nd._plan_sobel, vol.ndim, (axis, mode, cval,), {}))


def minimum_filter(input, size=None, footprint=None, mode="reflect", cval=0.0, origin=0, *, axes=None):
    vol = _check(input)
    return vol._run_plan(_nd._plan_minimum_filter(vol.ndim, size, footprint, mode, cval, origin, axes))


def maximum_filter(input, size=None, footprint=None, mode="reflect", cval=0.0, origin=0, *, axes=None):
    vol = _check(input)
    return vol._run_plan(_nd._plan_maximum_filter(vol.ndim, size, footprint, mode, cval, origin, axes))


FUNCTIONS = {f.__name__: f for f in (
    convolve, correlate, laplace, prewitt, sobel, gaussian_filter, gaussian_gradient_magnitude,
    gaussian_laplace, uniform_filter, minimum_filter, maximum_filter)}

assert _lib is not None
