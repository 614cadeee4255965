"""ctypes binding of ``libb200ndfilters.so`` (the C ABI in include/b200_ndfilters.h).

There is deliberately no fallback: if the CUDA library is missing or a call
fails, the caller gets an exception.  Error codes are translated to the Python
exception classes the scipy drivers raise for the same condition
(scipy/ndimage/_filters.py:600-608,1288-1292,1533-1536; _ni_support.py:59).
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# (B200_NDFILTERS_LIB selects another build of the same library: A/B runs of experimental kernels)
LIB_PATH = os.environ.get("B200_NDFILTERS_LIB") or os.path.join(_HERE, "libb200ndfilters.so")

# status codes / enums of include/b200_ndfilters.h
OK, EINVAL, EORIGIN, EMODE, EDTYPE, ECUDA, ENODEVICE = range(7)
FLAG_EXACT, FLAG_NO_TMA, FLAG_FORCE_FAST = 1, 2, 4
EPI_STORE, EPI_ACC, EPI_SQ_STORE, EPI_SQ_ACC, EPI_SQ_ACC_SQRT = range(5)
MAXDIM = 8

DTYPE_CODES = {
    "uint8": 0, "int8": 1, "uint16": 2, "int16": 3, "uint32": 4, "int32": 5,
    "uint64": 6, "int64": 7, "float32": 8, "float64": 9,
}

# scipy/ndimage/_ni_support.py:37-59 (filters map grid-wrap/grid-constant onto
# wrap/constant, probe-verified against scipy 1.18.1)
MODE_CODES = {
    "nearest": 0, "wrap": 1, "grid-wrap": 1, "reflect": 2, "grid-mirror": 2,
    "mirror": 3, "constant": 4, "grid-constant": 4,
}

EXPORTS = (
    "b200_correlate1d", "b200_uniform_filter1d", "b200_correlate_nd", "b200_combine",
    "b200_minmax_filter1d", "b200_minmax_filter_nd", "b200_correlate1d_halo", "b200_correlate1d_halo_eligible",
    "b200_enable_peer_access", "b200_uniform_filter1d_halo", "b200_uniform_filter1d_halo_eligible",
    "b200_correlate1d_pair", "b200_correlate1d_pair_eligible",
    "b200_uniform_filter1d_pair", "b200_uniform_filter1d_pair_eligible",
    "b200_ipc_export", "b200_ipc_open", "b200_ipc_close",
    "b200_correlate_nd_halo", "b200_correlate_nd_halo_eligible",
    "b200_shutdown", "b200_last_kernel", "b200_last_error", "b200_version", "b200_launch_count",
)

_lib = None


class B200Error(RuntimeError):
    """CUDA / driver level failure reported by the native library."""


def mode_code(mode):
    try:
        return MODE_CODES[mode]
    except (KeyError, TypeError):
        raise RuntimeError("boundary mode not supported")


def load():
    """Load the native library (once).  Fails loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "dask_image_b200: %s is missing -- build it with "
            "`python -m dask_image_b200.build` (nvcc, sm_100a); there is no CPU fallback"
            % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    vp, i64, dbl, ci, cu = (ctypes.c_void_p, ctypes.c_int64, ctypes.c_double, ctypes.c_int,
                            ctypes.c_uint)
    i64p = ctypes.POINTER(ctypes.c_int64)
    dblp = ctypes.POINTER(ctypes.c_double)
    lib.b200_correlate1d.argtypes = [vp, vp, ci, ci, i64p, ci, dblp, i64, ci, dbl, i64, i64, i64,
                                     ci, cu, vp]
    lib.b200_correlate1d.restype = ci
    lib.b200_uniform_filter1d.argtypes = [vp, vp, ci, ci, i64p, ci, i64, ci, dbl, i64, i64, i64,
                                          cu, vp]
    lib.b200_uniform_filter1d.restype = ci
    lib.b200_correlate_nd.argtypes = [vp, vp, ci, ci, i64p, dblp, i64p, i64p, ci, dbl, i64, i64,
                                      cu, vp]
    lib.b200_correlate_nd.restype = ci
    lib.b200_correlate1d_halo.argtypes = [vp, vp, ci, ci, i64p, dblp, i64, ci, dbl, i64, vp, i64, vp, i64, ci, cu, vp]
    lib.b200_correlate1d_halo.restype = ci
    lib.b200_uniform_filter1d_halo.argtypes = [vp, vp, ci, ci, i64p, i64, ci, dbl, i64, vp, i64, vp, i64, cu, vp]
    lib.b200_uniform_filter1d_halo.restype = ci
    lib.b200_uniform_filter1d_halo_eligible.argtypes = [ci, i64, i64, i64, i64, ci, dbl]
    lib.b200_uniform_filter1d_halo_eligible.restype = ci
    lib.b200_correlate1d_pair.argtypes = [vp, vp, ci, ci, i64p, dblp, i64, ci, i64, dblp, i64, ci, i64, dbl, ci, cu, vp]
    lib.b200_correlate1d_pair.restype = ci
    lib.b200_correlate1d_pair_eligible.argtypes = [ci, ci, i64p, i64, i64, i64, i64]
    lib.b200_correlate1d_pair_eligible.restype = ci
    lib.b200_uniform_filter1d_pair.argtypes = [vp, vp, ci, ci, i64p, i64, ci, i64, i64, ci, i64, dbl, cu, vp]
    lib.b200_uniform_filter1d_pair.restype = ci
    lib.b200_uniform_filter1d_pair_eligible.argtypes = [ci, ci, i64p, i64, i64, ci, i64, i64, ci, dbl]
    lib.b200_uniform_filter1d_pair_eligible.restype = ci
    lib.b200_correlate_nd_halo.argtypes = [vp, vp, ci, ci, i64p, dblp, i64p, i64p, ci, dbl, vp, i64, vp, i64, cu, vp]
    lib.b200_correlate_nd_halo.restype = ci
    lib.b200_correlate_nd_halo_eligible.argtypes = [ci, ci, i64p, i64p, i64p]
    lib.b200_correlate_nd_halo_eligible.restype = ci
    lib.b200_ipc_export.argtypes = [vp, vp, i64p]
    lib.b200_ipc_export.restype = ci
    lib.b200_ipc_open.argtypes = [vp, ci, ctypes.POINTER(vp)]
    lib.b200_ipc_open.restype = ci
    lib.b200_ipc_close.argtypes = [vp]
    lib.b200_ipc_close.restype = ci
    lib.b200_enable_peer_access.argtypes = [ci]
    lib.b200_enable_peer_access.restype = ci
    lib.b200_correlate1d_halo_eligible.argtypes = [ci, i64, i64, i64, i64]
    lib.b200_correlate1d_halo_eligible.restype = ci
    u8p = ctypes.POINTER(ctypes.c_uint8)
    lib.b200_minmax_filter1d.argtypes = [vp, vp, ci, ci, i64p, ci, i64, ci, ci, dbl, i64, i64, i64, cu, vp]
    lib.b200_minmax_filter1d.restype = ci
    lib.b200_minmax_filter_nd.argtypes = [vp, vp, ci, ci, i64p, u8p, i64p, i64p, ci, ci, dbl, i64, i64, cu, vp]
    lib.b200_minmax_filter_nd.restype = ci
    lib.b200_combine.argtypes = [vp, vp, ci, i64, ci, vp]
    lib.b200_combine.restype = ci
    for name in ("b200_last_kernel", "b200_last_error", "b200_version"):
        getattr(lib, name).argtypes = []
        getattr(lib, name).restype = ctypes.c_char_p
    lib.b200_shutdown.argtypes = []
    lib.b200_shutdown.restype = ci
    lib.b200_launch_count.argtypes = [ci]
    lib.b200_launch_count.restype = i64
    _lib = lib
    return lib


def last_kernel():
    return load().b200_last_kernel().decode()


def launch_count(reset=False):
    return int(load().b200_launch_count(1 if reset else 0))


def version():
    return load().b200_version().decode()


def check(rc):
    """Translate a non-zero status into the exception scipy would raise."""
    if rc == OK:
        return
    msg = load().b200_last_error().decode()
    if rc == EORIGIN:
        raise ValueError(msg)
    if rc in (EINVAL, EMODE, EDTYPE):
        raise RuntimeError(msg)
    raise B200Error(msg)


def i64_array(values):
    return (ctypes.c_int64 * len(values))(*[int(v) for v in values])


def f64_pointer(np_array):
    return np_array.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
