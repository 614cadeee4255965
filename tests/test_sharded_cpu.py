"""Slab engine on CPU: world_size 2 and 3 over ``gloo`` with the oracle as the
arithmetic backend.  Checks that partitioning + halo exchange + end-slab
boundary handling reproduce the whole-array oracle BIT FOR BIT (the reference's
own invariant: map_overlap(chunks) == whole-array scipy), for
every filter of the hot path, all boundary modes (wrap closes the ring) and
uneven slabs."""
import os
import socket
import sys
import traceback

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

MODES = ["reflect", "mirror", "nearest", "constant", "wrap"]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _cases():
    rng = np.random.default_rng(7)
    f32 = rng.random((23, 10, 12), dtype=np.float32)
    f64 = rng.random((19, 9, 8))
    u16 = rng.integers(0, 65536, (21, 8, 10), dtype=np.uint16)
    i32 = rng.integers(-1000, 1000, (17, 6), dtype=np.int32)
    w3 = rng.random((3, 4, 5)) - 0.4
    w2 = rng.random((5, 2)) - 0.4
    cases = []
    for mode in MODES:
        cases += [
            ("gaussian_filter", f32, dict(sigma=1.0, mode=mode, cval=0.3)),
            ("gaussian_filter", u16, dict(sigma=(1.5, 0.0, 0.8), order=(0, 0, 1), mode=mode)),
            ("uniform_filter", u16, dict(size=(5, 3, 4), mode=mode, origin=(1, 0, -2), cval=7.0)),
            ("uniform_filter", f64, dict(size=4, mode=mode)),
            ("gaussian_gradient_magnitude", f32, dict(sigma=1.2, mode=mode)),
            ("gaussian_laplace", f64, dict(sigma=(1.0, 0.7, 1.1), mode=mode, cval=-1.0)),
            ("correlate", f32, dict(weights=w3, mode=mode, origin=(1, -1, 2), cval=0.5)),
            ("convolve", i32, dict(weights=w2, mode=mode, cval=3.0)),
            ("sobel", f32, dict(axis=-1, mode=mode)),
            ("prewitt", u16, dict(axis=0, mode=mode)),
            ("laplace", f64, dict(mode=mode)),
            ("minimum_filter", u16, dict(size=(5, 1, 3), mode=mode, origin=(1, 0, -1), cval=9.0)),
            ("maximum_filter", f32, dict(footprint=np.array([[[1, 0, 1]], [[0, 1, 0]], [[1, 1, 0]]], bool),
                                         mode=mode, cval=0.4)),
            # scipy's `axes=` travels through **kwargs of the reference's wrappers (ref:_gaussian.py:119-123,147-151)
            ("gaussian_gradient_magnitude", u16, dict(sigma=1.0, mode=mode, axes=(2, 0))),
            # the pass along axis 0 comes AFTER another pass: it reads an intermediate, not the input
            ("gaussian_filter", u16, dict(sigma=(1.2, 0.8), mode=mode, axes=(2, 0))),
            ("uniform_filter", f32, dict(size=(3, 5), mode=mode, axes=(1, 0), origin=(0, 1))),
            # (the reference expands sigma to one entry per array axis, ref:_gaussian.py:22-44, which scipy's
            # gaussian_laplace only accepts together with ALL axes: a permutation changes the summation order)
            ("gaussian_laplace", f32, dict(sigma=(0.9, 1.3, 0.6), mode=mode, axes=(2, 0, 1), cval=0.2)),
        ]
    return cases


def _reference(name, a, kw):
    from oracle import oracle as orc
    import scipy.ndimage as ndi

    if hasattr(orc, name) and "axes" not in kw:
        return getattr(orc, name)(a, **kw)
    return getattr(ndi, name)(a, **kw)          # laplace / sobel / prewitt / `axes=`: scipy itself


def _worker(rank, world, port, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("gloo", rank=rank, world_size=world)
        from dask_image_b200 import ndfilters
        from dask_image_b200.sharded import ShardedVolume
        from _oracle_backend import OracleBackend

        failures = []
        for name, a, kw in _cases():
            vol = ShardedVolume.from_global(a, device="cpu", backend=OracleBackend, halo=2)
            if "axes" in kw and name in ("gaussian_filter", "uniform_filter"):
                # the reference's signatures of these two have no `axes`; the slab engine's functions do
                from dask_image_b200 import sharded

                res = sharded.FUNCTIONS[name](vol, **kw)
            else:
                res = getattr(ndfilters, name)(vol, **kw)
            assert isinstance(res, ShardedVolume) and res.shape == a.shape and res.dtype == a.dtype
            got = res.gather()
            ref = _reference(name, a, kw)
            if name == "uniform_filter" and a.dtype.kind == "f":
                # the reference's float box filter is a RUNNING sum along each
                # line (scipy/ndimage/_filters.py:1542): its rounding depends on
                # where the line starts, so chunked == whole only to rounding
                ok = np.allclose(got, ref, rtol=1e-12, atol=1e-12)
            else:
                ok = np.array_equal(got, ref, equal_nan=True)
            if not ok:
                failures.append("%s %s %s: %d voxels differ" % (
                    name, a.dtype, kw.get("mode"), int((got != ref).sum())))
        # synthetic volumes do not depend on the number of ranks
        for dt in (np.float32, np.uint16, np.int32):
            v = ShardedVolume.random((11, 5, 6), dt, seed=3, device="cpu", backend=OracleBackend)
            g = v.gather()
            one = torch.arange(g.size, dtype=torch.int64)
            from dask_image_b200.sharded import _hash_values, torch_dtype
            exp = _hash_values(one, 3, torch_dtype(dt))
            from _oracle_backend import _np
            if not np.array_equal(g.ravel(), _np(exp).ravel()):
                failures.append("random(%s) depends on the partition" % np.dtype(dt))
        # small volumes may be REPLICATED instead of cut (the default below REPLICATE_BELOW_BYTES; the test
        # session switches the rule off, so ask for it): every rank holds and filters the whole volume, no
        # exchange, same result -- even for a halo deeper than a slab would be
        rng = np.random.default_rng(5)
        a = rng.random((world * 2 + 1, 9, 10)).astype(np.float32)
        for kw in (dict(replicate=True), dict()):
            old_limit = ShardedVolume.REPLICATE_BELOW_BYTES
            if not kw:
                ShardedVolume.REPLICATE_BELOW_BYTES = 1 << 20       # the automatic rule
            try:
                vol = ShardedVolume.from_global(a, device="cpu", backend=OracleBackend, **kw)
            finally:
                ShardedVolume.REPLICATE_BELOW_BYTES = old_limit
            res = ndfilters.gaussian_filter(vol, sigma=3.0)
            if not (vol.replicated and res.replicated and vol.n_local == a.shape[0] and vol.world == 1):
                failures.append("small volume was not replicated (%s)" % (kw,))
            if not np.array_equal(res.gather(), _reference("gaussian_filter", a, dict(sigma=3.0))):
                failures.append("replicated result differs")
        # a halo deeper than the thinnest slab is refused on every rank
        vol = ShardedVolume.from_global(np.zeros((world * 2, 4), np.float32), device="cpu",
                                        backend=OracleBackend)
        try:
            ndfilters.gaussian_filter(vol, sigma=3.0)
            failures.append("deep halo accepted")
        except ValueError:
            pass
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, failures))
    except Exception:
        q.put((rank, ["EXC " + traceback.format_exc()]))


@pytest.mark.parametrize("world", [2, 3])
def test_slab_engine_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, failures in results:
        assert not failures, "rank %d: %s" % (rank, "\n".join(failures))


def test_slab_bounds():
    from dask_image_b200.sharded import slab_bounds

    assert slab_bounds(10, 3) == [(0, 4), (4, 7), (7, 10)]
    assert slab_bounds(2048, 8)[-1] == (1792, 2048)
    assert slab_bounds(2, 4) == [(0, 1), (1, 2), (2, 2), (2, 2)]


def test_single_rank_plan_on_cpu():
    """world == 1: no exchange, no pads -- the plan runs through the backend as
    a whole array and equals the oracle."""
    from dask_image_b200 import ndfilters
    from dask_image_b200.sharded import ShardedVolume
    from _oracle_backend import OracleBackend
    from oracle import oracle as orc

    a = np.random.default_rng(0).integers(0, 255, (9, 7, 5), dtype=np.uint8)
    vol = ShardedVolume.from_global(a, device="cpu", backend=OracleBackend)
    got = ndfilters.gaussian_gradient_magnitude(vol, sigma=1.0).gather()
    assert np.array_equal(got, orc.gaussian_gradient_magnitude(a, 1.0))
    got = ndfilters.uniform_filter(vol, size=3, mode="wrap").gather()
    assert np.array_equal(got, orc.uniform_filter(a, 3, mode="wrap"))
