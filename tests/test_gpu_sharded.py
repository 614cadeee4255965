"""GPU tests of the slab contract (pad_lo / pad_hi of the C ABI) and of the
NCCL slab engine.  The invariant under test is north_star's: the whole-volume
result is IDENTICAL (bit for bit) regardless of how many slabs / GPUs it was
computed on -- on top of parity with the oracle."""
import os
import socket
import sys
import traceback

import numpy as np
import pytest

from _helpers import assert_parity
from oracle import oracle as orc

pytestmark = pytest.mark.gpu

MODES = ["reflect", "mirror", "nearest", "constant", "wrap"]


@pytest.fixture(scope="module")
def b200():
    import dask_image_b200

    return dask_image_b200


def _slabs(n, parts):
    from dask_image_b200.sharded import slab_bounds

    return [b for b in slab_bounds(n, parts) if b[1] > b[0]]


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.uint16, np.int32])
def test_slab_calls_equal_whole_call(b200, dtype, mode):
    """Filtering plane ranges [s, e) of a resident volume with pad = (s, N - e)
    must reproduce the whole-array call exactly, for the axis-0 1-D pass, the
    box filter and the N-D correlate."""
    import torch
    from dask_image_b200 import ndimage as nd

    rng = np.random.default_rng(11)
    shape = (37, 12, 64)
    if np.dtype(dtype).kind == "f":
        a = (rng.random(shape) - 0.3).astype(dtype)
    else:
        a = rng.integers(0, 1000, shape).astype(dtype)
    x = torch.from_numpy(a).cuda() if a.dtype != np.uint16 else torch.from_numpy(a.view(np.int16)).cuda().view(torch.uint16)
    w1 = rng.random(9) - 0.2
    w3 = rng.random((4, 3, 5)) - 0.4
    dt = np.dtype(dtype)

    def run(kind, src, dst, n, pad):
        shp = (n,) + shape[1:]
        if kind == "corr":
            nd._launch_correlate1d(src, dst, shp, dt, 0, w1, mode, 0.7, 1, pad)
        elif kind == "uniform":
            nd._launch_uniform1d(src, dst, shp, dt, 0, 6, mode, 0.7, -1, pad)
        else:
            nd._launch_correlate_nd(src, dst, shp, dt, w3, (1, 0, -2), mode, 0.7, pad)

    for kind in ("corr", "uniform", "nd"):
        whole = torch.empty_like(x)
        run(kind, x, whole, shape[0], (0, 0))
        for parts in (2, 3, 5):
            out = torch.zeros_like(x)
            for s, e in _slabs(shape[0], parts):
                run(kind, x[s:e], out[s:e], e - s, (s, shape[0] - e))
            assert torch.equal(out.view(torch.uint8), whole.view(torch.uint8)), (kind, parts, mode, dt)
    # and the whole call itself is right
    got = b200.B200Array(whole).to_numpy()
    assert_parity(got, orc.correlate(a, w3, mode, 0.7, (1, 0, -2)), what="nd")


def test_sharded_single_rank_equals_array_path(b200):
    rng = np.random.default_rng(5)
    a = rng.random((40, 24, 64), dtype=np.float32)
    vol = b200.ShardedVolume.from_global(a)
    arr = b200.B200Array.from_numpy(a)
    for name, kw in [("gaussian_filter", dict(sigma=1.5)),
                     ("gaussian_gradient_magnitude", dict(sigma=1.0, mode="nearest")),
                     ("uniform_filter", dict(size=5, mode="wrap")),
                     ("convolve", dict(weights=rng.random((3, 3, 3)), mode="wrap"))]:
        f = getattr(b200.ndfilters, name)
        kw2 = dict(kw)
        if "weights" in kw2:
            kw2["weights"] = b200.B200Array.from_numpy(kw2["weights"])
        got = f(vol, **kw).gather()
        one = f(arr, **kw2).to_numpy()
        assert np.array_equal(got, one), name
        assert_parity(got, getattr(orc, name)(a, **kw), what=name)


def test_random_volume_is_partition_free(b200):
    import torch

    v = b200.ShardedVolume.random((16, 8, 32), np.float32, seed=9)
    g = v.gather()
    assert g.min() >= 0.0 and g.max() < 1.0 and abs(g.mean() - 0.5) < 0.05
    u = b200.ShardedVolume.random((16, 8, 32), np.uint16, seed=9).gather()
    assert u.dtype == np.uint16 and u.max() > 60000 and u.min() < 5000
    assert torch.cuda.is_available()


# ---------------------------------------------------------------------------
# NCCL: one process per GPU
# ---------------------------------------------------------------------------
def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _nccl_worker(rank, world, port, q):
    try:
        import torch
        import torch.distributed as dist

        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        torch.cuda.set_device(rank)
        import datetime

        # (a rank that fails must not leave the others waiting in a collective for ever)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank),
                                timeout=datetime.timedelta(seconds=240))
        import dask_image_b200 as b200

        rng = np.random.default_rng(21)
        # at least 12 planes per rank: the deepest filter below reaches 8
        f32 = rng.random((max(50, 12 * world + 5), 20, 64), dtype=np.float32)
        u16 = rng.integers(0, 65536, (max(45, 12 * world + 1), 16, 24), dtype=np.uint16)
        w3 = rng.random((3, 5, 4)) - 0.3
        failures = []
        cases = []
        for mode in MODES:
            cases += [
                ("gaussian_filter", f32, dict(sigma=2.0, mode=mode, cval=0.2)),
                ("gaussian_gradient_magnitude", f32, dict(sigma=1.5, mode=mode)),
                ("gaussian_laplace", f32, dict(sigma=1.0, mode=mode)),
                ("uniform_filter", u16, dict(size=7, mode=mode)),
                ("gaussian_filter", u16, dict(sigma=1.0, mode=mode)),
                ("convolve", f32, dict(weights=w3, mode=mode, origin=(0, 1, -1))),
                ("sobel", f32, dict(axis=0, mode=mode)),
            ]
        for name, a, kw in cases:
            f = getattr(b200.ndfilters, name)
            got = f(b200.ShardedVolume.from_global(a, halo=4), **kw).gather()
            kw1 = dict(kw)
            if "weights" in kw1:
                kw1["weights"] = b200.B200Array.from_numpy(kw1["weights"])
            one = f(b200.B200Array.from_numpy(a), **kw1).to_numpy()
            if not np.array_equal(got, one):
                failures.append("%s %s %s: %d voxels differ from the 1-GPU result" % (
                    name, a.dtype, kw["mode"], int((got != one).sum())))
        # fused halo path: slabs thick enough for the peer-TMA kernel read their
        # neighbours' planes in place (CUDA IPC); with it on or off the result is
        # the 1-GPU result, bit for bit
        from dask_image_b200 import ndimage as nd

        thick = rng.random((80 * world + 3, 8, 64), dtype=np.float32)
        sizes = {e - s for s, e in b200.sharded.slab_bounds(thick.shape[0], world)}
        assert all(nd._halo_eligible(np.dtype(np.float32), n, 8 * 64, 17, 0) for n in sizes)
        for mode in MODES:
            for name, kw in [("gaussian_filter", dict(sigma=2.0, mode=mode, cval=0.1)),
                             ("gaussian_gradient_magnitude", dict(sigma=(2.0, 1.0, 1.5), mode=mode)),
                             ("gaussian_laplace", dict(sigma=1.0, mode=mode))]:
                f = getattr(b200.ndfilters, name)
                one = f(b200.B200Array.from_numpy(thick), **kw).to_numpy()
                for flag in ("1", "0"):
                    os.environ["B200_PEER_HALO"] = flag
                    got = f(b200.ShardedVolume.from_global(thick, halo=8), **kw).gather()
                    if not np.array_equal(got, one):
                        failures.append("peer=%s %s %s: %d voxels differ from the 1-GPU result" % (
                            flag, name, mode, int((got != one).sum())))
        # N-D correlate / convolve with the neighbour planes read in place by the tiled kernel's first / last z tile
        from dask_image_b200 import _lib

        thick9 = rng.random((84 * world, 8, 64), dtype=np.float32)
        # (in-place halo reads need 8-row-aligned staged planes: 9 or 17 taps along axis 1; other weights take
        # the NCCL exchange -- checked as well)
        w9 = rng.random((9, 9, 5)) - 0.4
        w39 = rng.random((3, 9, 4)) - 0.4
        for arr, wts, org in ((thick9, w9, 0), (thick, w39, (0, 1, -1)), (thick9, w39, (-1, 0, 1)), (thick, w3, 0)):
            for mode in MODES:
                for name in ("convolve", "correlate"):
                    f = getattr(b200.ndfilters, name)
                    one = f(b200.B200Array.from_numpy(arr), b200.B200Array.from_numpy(wts), mode=mode, cval=0.3,
                            origin=org).to_numpy()
                    for flag in ("1", "0"):
                        os.environ["B200_PEER_HALO"] = flag
                        got = f(b200.ShardedVolume.from_global(arr, halo=8), wts, mode=mode, cval=0.3, origin=org).gather()
                        # (a rank with no neighbour plane to read -- an end slab whose reach towards its only
                        # neighbour is zero -- runs the plain kernel; with reach on both sides every rank reads)
                        both = wts.shape == (9, 9, 5)
                        want = "corrnd_tiled_peer_f32" if flag == "1" and both else "corrnd_tiled_f32"
                        if (both or flag == "0") and _lib.last_kernel() != want:
                            failures.append("peer=%s %s %s: kernel %s" % (flag, name, mode, _lib.last_kernel()))
                        if not np.array_equal(got, one):
                            failures.append("peer=%s %s %s %s: %d voxels differ from the 1-GPU result" % (
                                flag, name, wts.shape, mode, int((got != one).sum())))
        thick16 = rng.integers(0, 65536, (80 * world + 3, 8, 64), dtype=np.uint16)
        for mode in MODES:
            for arr in (thick16, thick):
                kw = dict(size=(7, 3, 5), mode=mode, origin=(1, 0, -1), cval=11.0)
                one = b200.ndfilters.uniform_filter(b200.B200Array.from_numpy(arr), **kw).to_numpy()
                for flag in ("1", "0"):
                    os.environ["B200_PEER_HALO"] = flag
                    got = b200.ndfilters.uniform_filter(b200.ShardedVolume.from_global(arr, halo=8), **kw).gather()
                    if not np.array_equal(got, one):
                        failures.append("peer=%s uniform_filter %s %s: %d voxels differ from the 1-GPU result" % (
                            flag, arr.dtype, mode, int((got != one).sum())))
        os.environ["B200_PEER_HALO"] = "1"
        # host-resident slabs streamed through each GPU (H2D | filter | D2H pipeline,
        # faces from the neighbours over NCCL): still bit-identical to one GPU
        from dask_image_b200.sharded import slab_bounds

        big = rng.random((30 * world + 3, 12, 64), dtype=np.float32)
        s0, s1 = slab_bounds(big.shape[0], world)[rank]
        pb = 12 * 64 * 4
        for mode in MODES:
            for name, kw in [("gaussian_filter", dict(sigma=2.0, mode=mode)),
                             ("gaussian_laplace", dict(sigma=1.0, mode=mode)),
                             ("uniform_filter", dict(size=(5, 3, 4), mode=mode, origin=(1, 0, -1))),
                             ("convolve", dict(weights=w3, mode=mode)),
                             ("sobel", dict(axis=1, mode=mode))]:
                f = getattr(b200.ndfilters, name)
                for chunk_planes in (8, 64):
                    hv = b200.HostVolume(big[s0:s1].copy(), global_shape=big.shape, chunk_bytes=chunk_planes * pb)
                    got = f(hv, **kw).to_numpy()
                    kw1 = dict(kw)
                    if "weights" in kw1:
                        kw1["weights"] = b200.B200Array.from_numpy(kw1["weights"])
                    one = f(b200.B200Array.from_numpy(big), **kw1).to_numpy()[s0:s1]
                    if not np.array_equal(got, one):
                        failures.append("HostVolume %s %s chunk=%d: %d voxels differ from the 1-GPU result" % (
                            name, mode, chunk_planes, int((got != one).sum())))
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, failures))
    except Exception:
        q.put((rank, ["EXC " + traceback.format_exc()]))


@pytest.mark.parametrize("world", [2, 4, 8])
def test_slab_engine_nccl(world):
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_nccl_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, failures in results:
        assert not failures, "rank %d: %s" % (rank, "\n".join(failures))


# ---------------------------------------------------------------------------
# HostVolume: host-resident volumes streamed through the GPU in axis-0 chunks
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("mode", MODES)
def test_host_volume_streaming_equals_resident(b200, mode):
    """Chunked H2D | filter | D2H pipeline == whole resident volume, bit for bit."""
    import torch

    rng = np.random.default_rng(31)
    a = rng.random((70, 24, 64), dtype=np.float32)
    u = rng.integers(0, 65536, (70, 16, 32), dtype=np.uint16)
    w3 = rng.random((5, 3, 3)) - 0.2
    plane = 24 * 64 * 4
    cases = [
        ("gaussian_filter", a, dict(sigma=1.5, mode=mode, cval=0.4)),
        ("gaussian_gradient_magnitude", a, dict(sigma=1.0, mode=mode)),
        ("gaussian_laplace", a, dict(sigma=(2.0, 1.0, 0.5), mode=mode)),
        ("uniform_filter", u, dict(size=(9, 3, 5), mode=mode, origin=(2, 0, -1))),
        ("convolve", a, dict(weights=w3, mode=mode, origin=(1, 0, 0))),
        ("sobel", a, dict(axis=0, mode=mode)),
        ("sobel", a, dict(axis=2, mode=mode)),          # axis-0 pass is not first: un-pipelined path
        ("laplace", a, dict(mode=mode)),
    ]
    for name, arr, kw in cases:
        f = getattr(b200.ndfilters, name)
        pb = int(np.prod(arr.shape[1:])) * arr.dtype.itemsize
        for chunk_planes in (7, 16, 64):
            hv = b200.HostVolume(arr, chunk_bytes=chunk_planes * pb)
            got = f(hv, **kw)
            assert isinstance(got, b200.HostVolume) and got.shape == arr.shape and got.dtype == arr.dtype
            kw1 = dict(kw)
            if "weights" in kw1:
                kw1["weights"] = b200.B200Array.from_numpy(kw1["weights"])
            one = f(b200.B200Array.from_numpy(arr), **kw1).to_numpy()
            assert np.array_equal(got.to_numpy(), one), (name, mode, chunk_planes)
    assert plane and torch.cuda.is_available()


def test_host_volume_pinned_output_reuse(b200):
    import torch

    a = np.random.default_rng(2).random((64, 32, 32), dtype=np.float32)
    h_in = torch.from_numpy(a).pin_memory()
    h_out = torch.empty_like(h_in).pin_memory()
    res = b200.ndfilters.gaussian_filter(b200.HostVolume(h_in, out=h_out, chunk_bytes=8 * 32 * 32 * 4), 2.0)
    assert res.tensor.data_ptr() == h_out.data_ptr()
    assert_parity(res.to_numpy(), orc.gaussian_filter(a, 2.0))


# ---------------------------------------------------------------------------
# b200_correlate1d_halo: neighbour planes read in place (the fused multi-GPU path)
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("mode", MODES)
def test_halo_in_place_equals_whole_call(b200, dtype, mode):
    """Slabs of one resident volume, their halos passed as pointers to the
    neighbouring planes (here: of the same volume; on several GPUs: a peer's
    memory) -- bit-identical to the whole-volume pass, volume boundaries included."""
    import torch
    from dask_image_b200 import _lib, ndimage as nd

    rng = np.random.default_rng(23)
    N, tail = 330, (6, 64)
    a = (rng.random((N,) + tail) - 0.4).astype(dtype)
    x = torch.from_numpy(a).cuda()
    dt = np.dtype(dtype)
    for nw, origin in ((17, 0), (9, 2), (33, -5), (4, 1), (1, 0)):
        w = rng.random(nw) - 0.3
        lo = nw // 2 + origin
        hi = nw - 1 - lo
        whole = torch.empty_like(x)
        nd._launch_correlate1d(x, whole, a.shape, dt, 0, w, mode, 0.3, origin)
        for bounds in ([(0, 160), (160, 330)], [(0, 100), (100, 215), (215, 330)]):
            if not all(nd._halo_eligible(dt, e - s, int(np.prod(tail)), nw, origin) for s, e in bounds):
                continue
            out = torch.zeros_like(x)
            for k, (s, e) in enumerate(bounds):
                ring = mode == "wrap"
                first, last = k == 0, k == len(bounds) - 1
                halo_lo = halo_hi = None
                if lo > 0 and (not first or ring):
                    halo_lo = x[s - lo:s] if not first else x[N - lo:N]
                if hi > 0 and (not last or ring):
                    halo_hi = x[e:e + hi] if not last else x[0:hi]
                nd._launch_correlate1d_halo(x[s:e], out[s:e], (e - s,) + tail, dt, w, mode, 0.3, origin,
                                            halo_lo, halo_hi)
                assert _lib.last_kernel().startswith("corr1d_tma_strided_peer"), _lib.last_kernel()
            assert torch.equal(out.view(torch.uint8), whole.view(torch.uint8)), (nw, origin, mode, bounds)
    # a slab too thin for the tile geometry is refused, not mis-computed
    assert not nd._halo_eligible(dt, 20, int(np.prod(tail)), 17, 0)
    with pytest.raises(RuntimeError, match="not eligible"):
        nd._launch_correlate1d_halo(x[:20], torch.empty_like(x[:20]), (20,) + tail, dt, np.ones(17), mode, 0.0, 0,
                                    None, x[20:28])


@pytest.mark.parametrize("dtype", [np.uint16, np.uint8, np.float32])
@pytest.mark.parametrize("mode", MODES)
def test_uniform_halo_in_place_equals_whole_call(b200, dtype, mode):
    import torch
    from dask_image_b200 import _lib, ndimage as nd

    rng = np.random.default_rng(29)
    N, tail = 300, (4, 256)
    if np.dtype(dtype).kind == "f":
        a = rng.random((N,) + tail).astype(dtype)
    else:
        a = rng.integers(0, np.iinfo(dtype).max, (N,) + tail, endpoint=True).astype(dtype)
    x = b200.B200Array.from_numpy(a).tensor
    dt = np.dtype(dtype)
    for size, origin in ((7, 0), (4, 1), (15, -7)):
        lo = size // 2 + origin
        hi = size - 1 - lo
        whole = torch.empty_like(x)
        nd._launch_uniform1d(x, whole, a.shape, dt, 0, size, mode, 5.0, origin)
        bounds = [(0, 140), (140, 300)]
        assert all(nd._uniform_halo_eligible(dt, e - s, int(np.prod(tail)), size, origin, mode, 5.0) for s, e in bounds)
        out = torch.zeros_like(x)
        for k, (s, e) in enumerate(bounds):
            ring = mode == "wrap"
            first, last = k == 0, k == len(bounds) - 1
            halo_lo = halo_hi = None
            if lo > 0 and (not first or ring):
                halo_lo = x[s - lo:s] if not first else x[N - lo:N]
            if hi > 0 and (not last or ring):
                halo_hi = x[e:e + hi] if not last else x[0:hi]
            nd._launch_uniform1d_halo(x[s:e], out[s:e], (e - s,) + tail, dt, size, mode, 5.0, origin, halo_lo, halo_hi)
            assert "_peer_" in _lib.last_kernel(), _lib.last_kernel()
        assert torch.equal(out.view(torch.uint8), whole.view(torch.uint8)), (size, origin, mode)


@pytest.mark.parametrize("mode", MODES)
def test_nd_halo_in_place_equals_whole_call(b200, mode):
    """b200_correlate_nd_halo: slabs of one resident volume whose neighbour planes are passed as pointers (on
    several GPUs: a peer's memory, read by the kernel's own TMA loads) -- bit-identical to the whole-volume call."""
    import torch
    from dask_image_b200 import _lib, ndimage as nd

    rng = np.random.default_rng(29)
    N, tail = 252, (24, 64)
    a = (rng.random((N,) + tail) - 0.4).astype(np.float32)
    x = torch.from_numpy(a).cuda()
    dt = np.dtype(np.float32)
    for wshape, origins in (((9, 9, 9), (0, 0, 0)), ((3, 9, 4), (0, 1, -1)), ((5, 9, 3), (-2, 0, 1)), ((4, 1, 7), (1, 0, 0))):
        w = rng.random(wshape) - 0.3
        lo = wshape[0] // 2 + origins[0]
        hi = wshape[0] - 1 - lo
        whole = torch.empty_like(x)
        nd._launch_correlate_nd(x, whole, a.shape, dt, w, origins, mode, 0.3)
        for bounds in ([(0, 124), (124, 252)], [(0, 84), (84, 168), (168, 252)]):
            assert all(nd._nd_halo_eligible(dt, (e - s,) + tail, wshape, origins) for s, e in bounds), (wshape, bounds)
            out = torch.zeros_like(x)
            for k, (s, e) in enumerate(bounds):
                ring = mode == "wrap"
                first, last = k == 0, k == len(bounds) - 1
                halo_lo = halo_hi = None
                if lo > 0 and (not first or ring):
                    halo_lo = x[s - lo:s] if not first else x[N - lo:N]
                if hi > 0 and (not last or ring):
                    halo_hi = x[e:e + hi] if not last else x[0:hi]
                nd._launch_correlate_nd_halo(x[s:e], out[s:e], (e - s,) + tail, dt, w, origins, mode, 0.3,
                                             halo_lo, halo_hi)
                if halo_lo is not None or halo_hi is not None:
                    assert _lib.last_kernel() == "corrnd_tiled_peer_f32", _lib.last_kernel()
            torch.cuda.synchronize()
            assert torch.equal(out, whole), "%s %s %s: %d voxels differ" % (
                wshape, bounds, mode, int((out != whole).sum().item()))
    # weights whose staged planes are not 8-row aligned are refused (the slab engine then exchanges the planes)
    assert not nd._nd_halo_eligible(dt, (84,) + tail, (3, 5, 4), (0, 0, 0))
    assert not nd._nd_halo_eligible(np.dtype(np.float64), (84,) + tail, (9, 9, 9), (0, 0, 0))


@pytest.mark.parametrize("mode", MODES)
def test_host_volume_out_of_core_equals_resident(b200, mode):
    """A host volume that does not fit the device memory it is allowed (max_device_bytes) is filtered in axis-0
    blocks that travel with their halo planes through two device windows: the result is the resident one, bit for
    bit -- block seams, volume boundaries and the periodic ring (mode='wrap') included."""
    rng = np.random.default_rng(31)
    a = (rng.random((90, 24, 64)) - 0.3).astype(np.float32)
    u = rng.integers(0, 65536, (75, 16, 32), dtype=np.uint16)
    w3 = rng.random((3, 5, 4)) - 0.3
    for arr in (a, u):
        plane = int(np.prod(arr.shape[1:])) * arr.dtype.itemsize
        cases = [("gaussian_filter", dict(sigma=2.0, mode=mode, cval=0.2)),
                 ("uniform_filter", dict(size=(5, 3, 4), mode=mode, origin=(1, 0, -1))),
                 ("gaussian_laplace", dict(sigma=1.0, mode=mode)),
                 ("gaussian_gradient_magnitude", dict(sigma=1.5, mode=mode)),
                 ("convolve", dict(weights=w3, mode=mode, cval=0.4, origin=(1, 0, -1)))]
        for name, kw in cases:
            f = getattr(b200.ndfilters, name)
            kw1 = dict(kw)
            if "weights" in kw1:
                kw1["weights"] = b200.B200Array.from_numpy(kw1["weights"])
            one = f(b200.B200Array.from_numpy(arr), **kw1).to_numpy()
            for budget_planes in (150, 400):          # ~10 blocks / ~3 blocks
                hv = b200.HostVolume(arr.copy(), max_device_bytes=budget_planes * plane)
                got = f(hv, **kw).to_numpy()
                assert np.array_equal(got, one), "%s %s %s budget=%d: %d voxels differ" % (
                    name, arr.dtype, mode, budget_planes, int((got != one).sum()))
    # a budget that cannot hold two blocks and their halos is refused
    with pytest.raises(MemoryError):
        b200.ndfilters.gaussian_filter(b200.HostVolume(a.copy(), max_device_bytes=20 * 24 * 64 * 4), 2.0)
