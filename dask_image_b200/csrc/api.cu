// api.cu -- extern "C" entry points declared in include/b200_ndfilters.h:
// argument validation (same conditions and error classes as the scipy drivers
// the reference calls), kernel-family selection, thread-local diagnostics.
#include "common.cuh"
#include <cuda.h>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <atomic>
#include <mutex>
#include <vector>

namespace b200 {

static thread_local char tl_error[512] = "";
static thread_local char tl_kernel[128] = "";
static thread_local int64_t tl_launches = 0;

int set_error(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(tl_error, sizeof(tl_error), fmt, ap);
    va_end(ap);
    return code;
}

void note_kernel(const char *name)
{
    strncpy(tl_kernel, name, sizeof(tl_kernel) - 1);
    tl_kernel[sizeof(tl_kernel) - 1] = 0;
}

void count_launch(int n) { tl_launches += n; }

int check_cuda(cudaError_t e, const char *what)
{
    if (e == cudaSuccess) return B200_OK;
    return set_error(B200_ECUDA, "CUDA error in %s: %s (%s)", what, cudaGetErrorName(e),
                     cudaGetErrorString(e));
}

int device_sm_count()
{
    static std::atomic<int> cached[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    int v = cached[dev].load(std::memory_order_relaxed);
    if (v == 0) {
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        cached[dev].store(v, std::memory_order_relaxed);
    }
    return v;
}

namespace {
struct TableEntry {
    int dev;
    uint64_t hash, stamp;
    std::vector<unsigned char> copy;
    void *dptr;
};
std::mutex g_table_mu;
std::vector<TableEntry> g_tables;
uint64_t g_table_stamp = 0;
}  // namespace

const void *cached_device_table(const void *host, size_t bytes)
{
    int dev = 0;
    if (check_cuda(cudaGetDevice(&dev), "cudaGetDevice")) return nullptr;
    uint64_t h = 1469598103934665603ull;
    const unsigned char *b = (const unsigned char *)host;
    for (size_t i = 0; i < bytes; ++i) h = (h ^ b[i]) * 1099511628211ull;
    std::lock_guard<std::mutex> lk(g_table_mu);
    for (auto &e : g_tables)
        if (e.dev == dev && e.hash == h && e.copy.size() == bytes && memcmp(e.copy.data(), host, bytes) == 0) {
            e.stamp = ++g_table_stamp;
            return e.dptr;
        }
    if (g_tables.size() >= 32) {
        size_t lru = 0;
        for (size_t i = 1; i < g_tables.size(); ++i)
            if (g_tables[i].stamp < g_tables[lru].stamp) lru = i;
        cudaFree(g_tables[lru].dptr);      // (synchronises the device: no launch still reads the table)
        g_tables.erase(g_tables.begin() + lru);
    }
    void *d = nullptr;
    if (check_cuda(cudaMalloc(&d, bytes ? bytes : 1), "cudaMalloc (filter table)")) return nullptr;
    if (check_cuda(cudaMemcpy(d, host, bytes, cudaMemcpyHostToDevice), "cudaMemcpy (filter table)")) {
        cudaFree(d);
        return nullptr;
    }
    TableEntry e;
    e.dev = dev; e.hash = h; e.stamp = ++g_table_stamp; e.dptr = d;
    e.copy.assign(b, b + bytes);
    g_tables.push_back(std::move(e));
    return d;
}

static int check_common(const void *in, const void *out, int dtype, int ndim, const int64_t *shape,
                        int mode, int64_t *total)
{
    if (dtype < B200_U8 || dtype > B200_F64) return set_error(B200_EDTYPE, "data type not supported");
    if (ndim < 1 || ndim > B200_MAXDIM) return set_error(B200_EINVAL, "ndim must be in [1, %d]", B200_MAXDIM);
    if (mode < B200_MODE_NEAREST || mode > B200_MODE_CONSTANT)
        return set_error(B200_EMODE, "boundary mode not supported");
    int64_t t = 1;
    for (int d = 0; d < ndim; ++d) {
        if (shape[d] < 0) return set_error(B200_EINVAL, "negative extent");
        t *= shape[d];
    }
    *total = t;
    if (t > 0 && (!in || !out)) return set_error(B200_EINVAL, "null device pointer");
    if (t > 0 && in == out) return set_error(B200_EINVAL, "input and output must not alias");
    return B200_OK;
}

}  // namespace b200

using namespace b200;

extern "C" {

int b200_correlate1d(const void *in, void *out, int dtype, int ndim, const int64_t *shape, int axis,
                     const double *weights, int64_t nw, int mode, double cval, int64_t origin,
                     int64_t pad_lo, int64_t pad_hi, int epilogue, unsigned flags, void *stream)
{
    int64_t total;
    if (int rc = check_common(in, out, dtype, ndim, shape, mode, &total)) return rc;
    if (axis < 0 || axis >= ndim) return set_error(B200_EINVAL, "axis out of range");
    if (!weights || nw < 1) return set_error(B200_EINVAL, "no filter weights given");
    if (nw > (1 << 20)) return set_error(B200_EINVAL, "too many filter weights");
    if (origin < -(nw / 2) || origin > (nw - 1) / 2)
        return set_error(B200_EORIGIN, "Invalid origin; origin must satisfy "
                                       "-(len(weights) // 2) <= origin <= (len(weights)-1) // 2");
    if (epilogue < B200_EPI_STORE || epilogue > B200_EPI_SQ_ACC_SQRT)
        return set_error(B200_EINVAL, "unknown epilogue");
    if (pad_lo < 0 || pad_hi < 0 || ((pad_lo || pad_hi) && axis != 0))
        return set_error(B200_EINVAL, "halo pads are only defined for axis 0");
    if (total == 0) return B200_OK;

    Corr1dArgs a;
    a.in = in; a.out = out; a.dtype = dtype;
    a.g = line_geometry(ndim, shape, axis);
    a.w = weights; a.nw = (int)nw; a.mode = mode; a.cval = cval; a.origin = (int)origin;
    a.pad_lo = pad_lo; a.pad_hi = pad_hi; a.epilogue = epilogue; a.flags = flags;
    a.stream = (cudaStream_t)stream;

    const bool is_float = dtype == B200_F32 || dtype == B200_F64;
    if (is_float && !(flags & (B200_FLAG_EXACT | B200_FLAG_NO_TMA))) {
        int rc = launch_corr1d_tma(a);
        if (rc >= 0) return rc;
    }
    if (!is_float && !(flags & B200_FLAG_EXACT)) {
        int rc = launch_corr1d_int(a);
        if (rc >= 0) return rc;
    }
    if (flags & B200_FLAG_FORCE_FAST)
        return set_error(B200_EINVAL, "no tiled kernel is eligible for this call");
    return launch_corr1d_generic(a);
}

int b200_correlate1d_halo(const void *in, void *out, int dtype, int ndim, const int64_t *shape,
                          const double *weights, int64_t nw, int mode, double cval, int64_t origin,
                          const void *halo_lo, int64_t n_lo, const void *halo_hi, int64_t n_hi, int epilogue,
                          unsigned flags, void *stream)
{
    int64_t total;
    if (int rc = check_common(in, out, dtype, ndim, shape, mode, &total)) return rc;
    if (!weights || nw < 1) return set_error(B200_EINVAL, "no filter weights given");
    if (nw > (1 << 20)) return set_error(B200_EINVAL, "too many filter weights");
    if (origin < -(nw / 2) || origin > (nw - 1) / 2)
        return set_error(B200_EORIGIN, "Invalid origin; origin must satisfy "
                                       "-(len(weights) // 2) <= origin <= (len(weights)-1) // 2");
    if (epilogue < B200_EPI_STORE || epilogue > B200_EPI_SQ_ACC_SQRT)
        return set_error(B200_EINVAL, "unknown epilogue");
    if (n_lo < 0 || n_hi < 0 || (n_lo > 0 && !halo_lo) || (n_hi > 0 && !halo_hi))
        return set_error(B200_EINVAL, "halo planes and their counts do not agree");
    if (total == 0) return B200_OK;
    if (dtype != B200_F32 && dtype != B200_F64)
        return set_error(B200_EINVAL, "not eligible: in-place halo reads are implemented for float32 / float64");
    Corr1dArgs a;
    a.in = in; a.out = out; a.dtype = dtype;
    a.g = line_geometry(ndim, shape, 0);
    a.w = weights; a.nw = (int)nw; a.mode = mode; a.cval = cval; a.origin = (int)origin;
    a.pad_lo = 0; a.pad_hi = 0; a.epilogue = epilogue; a.flags = flags;
    a.stream = (cudaStream_t)stream;
    a.peer_lo = n_lo > 0 ? halo_lo : nullptr; a.peer_hi = n_hi > 0 ? halo_hi : nullptr;
    a.n_peer_lo = (int)n_lo; a.n_peer_hi = (int)n_hi;
    int rc = launch_corr1d_peer(a);
    if (rc < 0) return set_error(B200_EINVAL, "not eligible: tile geometry cannot place the neighbour rows");
    return rc;
}

int b200_correlate1d_pair_eligible(int dtype, int ndim, const int64_t *shape, int64_t nw1, int64_t origin1,
                                   int64_t nw2, int64_t origin2)
{
    if (ndim < 2 || ndim > B200_MAXDIM || !shape) return 0;
    if (nw1 < 1 || nw1 > 4096 || nw2 < 1 || nw2 > 4096) return 0;
    return corr2d_fused_eligible(dtype, shape[ndim - 2], shape[ndim - 1], (int)nw1, (int)origin1, (int)nw2,
                                 (int)origin2);
}

int b200_correlate1d_pair(const void *in, void *out, int dtype, int ndim, const int64_t *shape, const double *w1,
                          int64_t nw1, int mode1, int64_t origin1, const double *w2, int64_t nw2, int mode2,
                          int64_t origin2, double cval, int epilogue, unsigned flags, void *stream)
{
    int64_t total;
    if (int rc = check_common(in, out, dtype, ndim, shape, mode1, &total)) return rc;
    if (mode2 < B200_MODE_NEAREST || mode2 > B200_MODE_CONSTANT)
        return set_error(B200_EMODE, "boundary mode not supported");
    if (ndim < 2) return set_error(B200_EINVAL, "a pass pair needs at least two axes");
    if (!w1 || nw1 < 1 || !w2 || nw2 < 1) return set_error(B200_EINVAL, "no filter weights given");
    if (origin1 < -(nw1 / 2) || origin1 > (nw1 - 1) / 2 || origin2 < -(nw2 / 2) || origin2 > (nw2 - 1) / 2)
        return set_error(B200_EORIGIN, "Invalid origin; origin must satisfy "
                                       "-(len(weights) // 2) <= origin <= (len(weights)-1) // 2");
    if (epilogue < B200_EPI_STORE || epilogue > B200_EPI_SQ_ACC_SQRT)
        return set_error(B200_EINVAL, "unknown epilogue");
    if (total == 0) return B200_OK;
    if (flags & B200_FLAG_EXACT) return set_error(B200_EINVAL, "not eligible: the fused pass pair is a tiled kernel");
    if (!b200_correlate1d_pair_eligible(dtype, ndim, shape, nw1, origin1, nw2, origin2))
        return set_error(B200_EINVAL, "not eligible: no fused kernel for this pair of filters");
    Corr2dArgs a;
    a.in = in; a.out = out; a.dtype = dtype;
    a.outer = 1;
    for (int d = 0; d < ndim - 2; ++d) a.outer *= shape[d];
    a.ny = shape[ndim - 2]; a.nx = shape[ndim - 1];
    a.w1 = w1; a.nw1 = (int)nw1; a.origin1 = (int)origin1; a.mode1 = mode1;
    a.w2 = w2; a.nw2 = (int)nw2; a.origin2 = (int)origin2; a.mode2 = mode2;
    a.cval = cval; a.epilogue = epilogue; a.flags = flags; a.stream = (cudaStream_t)stream;
    int rc = launch_corr2d_fused(a);
    if (rc < 0) return set_error(B200_EINVAL, "not eligible: no fused kernel for this pair of filters");
    return rc;
}

int b200_enable_peer_access(int peer_device)
{
    int dev = -1;
    B200_CUDA_TRY(cudaGetDevice(&dev));
    if (dev == peer_device) return B200_OK;
    int can = 0;
    B200_CUDA_TRY(cudaDeviceCanAccessPeer(&can, dev, peer_device));
    if (!can) return set_error(B200_ENODEVICE, "device %d cannot access device %d", dev, peer_device);
    cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) {
        cudaGetLastError();   // clear the sticky-less error state
        return B200_OK;
    }
    return check_cuda(e, "cudaDeviceEnablePeerAccess");
}

// ---- CUDA IPC: neighbour slabs mapped for the in-place halo reads ------------
typedef CUresult (*PFN_memGetAddressRange)(CUdeviceptr *, size_t *, CUdeviceptr);

int b200_ipc_export(const void *dev_ptr, void *handle_out, int64_t *offset_out)
{
    if (!dev_ptr || !handle_out || !offset_out) return set_error(B200_EINVAL, "null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == B200_IPC_HANDLE_BYTES, "IPC handle size");
    static PFN_memGetAddressRange get_range = nullptr;
    static std::atomic<int> looked{0};
    if (!looked.load()) {
        void *fp = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &fp, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            get_range = (PFN_memGetAddressRange)fp;
        looked.store(1);
    }
    if (!get_range) return set_error(B200_ECUDA, "cuMemGetAddressRange is not available");
    CUdeviceptr base = 0;
    size_t size = 0;
    if (get_range(&base, &size, (CUdeviceptr)(uintptr_t)dev_ptr) != CUDA_SUCCESS)
        return set_error(B200_ECUDA, "cuMemGetAddressRange failed for %p", dev_ptr);
    cudaIpcMemHandle_t h;
    B200_CUDA_TRY(cudaIpcGetMemHandle(&h, (void *)(uintptr_t)base));
    memcpy(handle_out, &h, sizeof(h));
    *offset_out = (int64_t)((uintptr_t)dev_ptr - (uintptr_t)base);
    return B200_OK;
}

int b200_ipc_open(const void *handle, int owner_device, void **base_out)
{
    if (!handle || !base_out) return set_error(B200_EINVAL, "null argument");
    int dev = -1;
    B200_CUDA_TRY(cudaGetDevice(&dev));
    if (owner_device >= 0 && owner_device != dev) {
        if (int rc = b200_enable_peer_access(owner_device)) return rc;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    void *p = nullptr;
    B200_CUDA_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *base_out = p;
    return B200_OK;
}

int b200_ipc_close(void *base)
{
    if (!base) return B200_OK;
    B200_CUDA_TRY(cudaIpcCloseMemHandle(base));
    return B200_OK;
}

int b200_correlate1d_halo_eligible(int dtype, int64_t n, int64_t inner, int64_t nw, int64_t origin)
{
    if (nw < 1 || nw > 4096) return 0;
    return corr1d_peer_eligible(dtype, n, inner, (int)nw, (int)origin);
}

int b200_uniform_filter1d(const void *in, void *out, int dtype, int ndim, const int64_t *shape, int axis,
                          int64_t size, int mode, double cval, int64_t origin, int64_t pad_lo,
                          int64_t pad_hi, unsigned flags, void *stream)
{
    int64_t total;
    if (int rc = check_common(in, out, dtype, ndim, shape, mode, &total)) return rc;
    if (axis < 0 || axis >= ndim) return set_error(B200_EINVAL, "axis out of range");
    if (size < 1 || size > (1 << 20)) return set_error(B200_EINVAL, "incorrect filter size");
    if (size / 2 + origin < 0 || size / 2 + origin >= size) return set_error(B200_EORIGIN, "invalid origin");
    if (pad_lo < 0 || pad_hi < 0 || ((pad_lo || pad_hi) && axis != 0))
        return set_error(B200_EINVAL, "halo pads are only defined for axis 0");
    if (total == 0) return B200_OK;

    Uniform1dArgs a;
    a.in = in; a.out = out; a.dtype = dtype;
    a.g = line_geometry(ndim, shape, axis);
    a.size = (int)size; a.mode = mode; a.cval = cval; a.origin = (int)origin;
    a.pad_lo = pad_lo; a.pad_hi = pad_hi; a.flags = flags; a.stream = (cudaStream_t)stream;
    if (!(flags & (B200_FLAG_EXACT | B200_FLAG_NO_TMA))) {
        int rc = launch_uniform1d_fast(a);
        if (rc >= 0) return rc;
    }
    if (flags & B200_FLAG_FORCE_FAST)
        return set_error(B200_EINVAL, "no tiled kernel is eligible for this call");
    return launch_uniform1d_generic(a);
}

int b200_uniform_filter1d_halo(const void *in, void *out, int dtype, int ndim, const int64_t *shape, int64_t size,
                               int mode, double cval, int64_t origin, const void *halo_lo, int64_t n_lo,
                               const void *halo_hi, int64_t n_hi, unsigned flags, void *stream)
{
    int64_t total;
    if (int rc = check_common(in, out, dtype, ndim, shape, mode, &total)) return rc;
    if (size < 1 || size > (1 << 20)) return set_error(B200_EINVAL, "incorrect filter size");
    if (size / 2 + origin < 0 || size / 2 + origin >= size) return set_error(B200_EORIGIN, "invalid origin");
    if (n_lo < 0 || n_hi < 0 || (n_lo > 0 && !halo_lo) || (n_hi > 0 && !halo_hi))
        return set_error(B200_EINVAL, "halo planes and their counts do not agree");
    if (total == 0) return B200_OK;
    Uniform1dArgs a;
    a.in = in; a.out = out; a.dtype = dtype;
    a.g = line_geometry(ndim, shape, 0);
    a.size = (int)size; a.mode = mode; a.cval = cval; a.origin = (int)origin;
    a.pad_lo = 0; a.pad_hi = 0; a.flags = flags; a.stream = (cudaStream_t)stream;
    int rc = launch_uniform1d_peer(a, n_lo > 0 ? halo_lo : nullptr, (int)n_lo, n_hi > 0 ? halo_hi : nullptr, (int)n_hi);
    if (rc < 0) return set_error(B200_EINVAL, "not eligible: no in-place halo kernel for this box filter");
    return rc;
}

int b200_uniform_filter1d_halo_eligible(int dtype, int64_t n, int64_t inner, int64_t size, int64_t origin, int mode,
                                        double cval)
{
    if (size < 1 || size > 4096) return 0;
    return uniform1d_peer_eligible(dtype, n, inner, (int)size, (int)origin, mode, cval);
}

int b200_uniform_filter1d_pair_eligible(int dtype, int ndim, const int64_t *shape, int64_t size1, int64_t origin1,
                                        int mode1, int64_t size2, int64_t origin2, int mode2, double cval)
{
    if (ndim < 2 || ndim > B200_MAXDIM || !shape) return 0;
    if (size1 < 1 || size1 > 4096 || size2 < 1 || size2 > 4096) return 0;
    return uniform2d_fused_eligible(dtype, shape[ndim - 2], shape[ndim - 1], (int)size1, (int)origin1, mode1,
                                    (int)size2, (int)origin2, mode2, cval);
}

int b200_uniform_filter1d_pair(const void *in, void *out, int dtype, int ndim, const int64_t *shape, int64_t size1,
                               int mode1, int64_t origin1, int64_t size2, int mode2, int64_t origin2, double cval,
                               unsigned flags, void *stream)
{
    int64_t total;
    if (int rc = check_common(in, out, dtype, ndim, shape, mode1, &total)) return rc;
    if (mode2 < B200_MODE_NEAREST || mode2 > B200_MODE_CONSTANT)
        return set_error(B200_EMODE, "boundary mode not supported");
    if (ndim < 2) return set_error(B200_EINVAL, "a pass pair needs at least two axes");
    if (size1 < 1 || size1 > (1 << 20) || size2 < 1 || size2 > (1 << 20))
        return set_error(B200_EINVAL, "incorrect filter size");
    if (size1 / 2 + origin1 < 0 || size1 / 2 + origin1 >= size1 || size2 / 2 + origin2 < 0 ||
        size2 / 2 + origin2 >= size2)
        return set_error(B200_EORIGIN, "invalid origin");
    if (total == 0) return B200_OK;
    if (flags & (B200_FLAG_EXACT | B200_FLAG_NO_TMA))
        return set_error(B200_EINVAL, "not eligible: the fused box-filter pair is a TMA-tiled kernel");
    Uniform2dArgs a;
    a.in = in; a.out = out; a.dtype = dtype;
    a.outer = 1;
    for (int d = 0; d < ndim - 2; ++d) a.outer *= shape[d];
    a.ny = shape[ndim - 2]; a.nx = shape[ndim - 1];
    a.size1 = (int)size1; a.origin1 = (int)origin1; a.mode1 = mode1;
    a.size2 = (int)size2; a.origin2 = (int)origin2; a.mode2 = mode2;
    a.cval = cval; a.flags = flags; a.stream = (cudaStream_t)stream;
    int rc = launch_uniform2d_fused(a);
    if (rc < 0)
        return set_error(B200_EINVAL, "not eligible: no fused kernel for this pair of box filters "
                                      "(u8 / u16, equal odd sizes, origin 0, 16-byte rows and pointers)");
    return rc;
}

int b200_correlate_nd(const void *in, void *out, int dtype, int ndim, const int64_t *shape,
                      const double *weights, const int64_t *wshape, const int64_t *origins, int mode,
                      double cval, int64_t pad_lo, int64_t pad_hi, unsigned flags, void *stream)
{
    int64_t total;
    if (int rc = check_common(in, out, dtype, ndim, shape, mode, &total)) return rc;
    if (!weights || !wshape || !origins) return set_error(B200_EINVAL, "no filter weights given");
    int64_t wtotal = 1;
    for (int d = 0; d < ndim; ++d) {
        if (wshape[d] < 1) return set_error(B200_EINVAL, "weights must have non-zero extents");
        wtotal *= wshape[d];
        if (wtotal > (1 << 24)) return set_error(B200_EINVAL, "too many filter weights");
        if (origins[d] < -(wshape[d] / 2) || origins[d] > (wshape[d] - 1) / 2)
            return set_error(B200_EORIGIN, "Invalid origin; origin must satisfy "
                                           "-(weights.shape[k] // 2) <= origin[k] <= (weights.shape[k]-1) // 2");
    }
    if (pad_lo < 0 || pad_hi < 0) return set_error(B200_EINVAL, "negative halo pad");
    if (total == 0) return B200_OK;

    CorrNdArgs a;
    a.in = in; a.out = out; a.dtype = dtype; a.ndim = ndim;
    for (int d = 0; d < ndim; ++d) {
        a.shape[d] = shape[d]; a.wshape[d] = wshape[d]; a.origins[d] = origins[d];
    }
    a.w = weights; a.mode = mode; a.cval = cval; a.pad_lo = pad_lo; a.pad_hi = pad_hi;
    a.flags = flags; a.stream = (cudaStream_t)stream;
    const bool is_float = dtype == B200_F32 || dtype == B200_F64;
    if (is_float && !(flags & (B200_FLAG_EXACT | B200_FLAG_NO_TMA))) {
        int rc = launch_corrnd_tiled(a);
        if (rc >= 0) return rc;
    }
    if (!is_float && !(flags & B200_FLAG_EXACT)) {
        int rc = launch_corrnd_int(a);
        if (rc >= 0) return rc;
    }
    if (flags & B200_FLAG_FORCE_FAST)
        return set_error(B200_EINVAL, "no tiled kernel is eligible for this call");
    return launch_corrnd_generic(a);
}

int b200_correlate_nd_halo_eligible(int dtype, int ndim, const int64_t *shape, const int64_t *wshape,
                                    const int64_t *origins)
{
    if (ndim != 3 || !shape || !wshape || !origins) return 0;
    for (int d = 0; d < ndim; ++d) {
        if (shape[d] < 1 || wshape[d] < 1) return 0;
        if (origins[d] < -(wshape[d] / 2) || origins[d] > (wshape[d] - 1) / 2) return 0;
    }
    return corrnd_peer_eligible(dtype, ndim, shape, wshape, origins);
}

int b200_correlate_nd_halo(const void *in, void *out, int dtype, int ndim, const int64_t *shape,
                           const double *weights, const int64_t *wshape, const int64_t *origins, int mode,
                           double cval, const void *halo_lo, int64_t n_lo, const void *halo_hi, int64_t n_hi,
                           unsigned flags, void *stream)
{
    int64_t total;
    if (int rc = check_common(in, out, dtype, ndim, shape, mode, &total)) return rc;
    if (!weights || !wshape || !origins) return set_error(B200_EINVAL, "no filter weights given");
    for (int d = 0; d < ndim; ++d) {
        if (wshape[d] < 1) return set_error(B200_EINVAL, "weights must have non-zero extents");
        if (origins[d] < -(wshape[d] / 2) || origins[d] > (wshape[d] - 1) / 2)
            return set_error(B200_EORIGIN, "Invalid origin; origin must satisfy "
                                           "-(weights.shape[k] // 2) <= origin[k] <= (weights.shape[k]-1) // 2");
    }
    if (n_lo < 0 || n_hi < 0 || (n_lo > 0 && !halo_lo) || (n_hi > 0 && !halo_hi))
        return set_error(B200_EINVAL, "halo planes and their counts do not agree");
    if (total == 0) return B200_OK;
    if (!b200_correlate_nd_halo_eligible(dtype, ndim, shape, wshape, origins))
        return set_error(B200_EINVAL, "not eligible: no in-place halo kernel for this N-D correlate");
    CorrNdArgs a;
    a.in = in; a.out = out; a.dtype = dtype; a.ndim = ndim;
    for (int d = 0; d < ndim; ++d) {
        a.shape[d] = shape[d]; a.wshape[d] = wshape[d]; a.origins[d] = origins[d];
    }
    a.w = weights; a.mode = mode; a.cval = cval; a.pad_lo = 0; a.pad_hi = 0;
    a.flags = flags; a.stream = (cudaStream_t)stream;
    a.peer_lo = n_lo > 0 ? halo_lo : nullptr; a.peer_hi = n_hi > 0 ? halo_hi : nullptr;
    a.n_peer_lo = (int)n_lo; a.n_peer_hi = (int)n_hi;
    int rc = launch_corrnd_peer(a);
    if (rc < 0) return set_error(B200_EINVAL, "not eligible: tile geometry cannot place the neighbour planes");
    return rc;
}

int b200_minmax_filter1d(const void *in, void *out, int dtype, int ndim, const int64_t *shape, int axis,
                         int64_t size, int is_max, int mode, double cval, int64_t origin, int64_t pad_lo,
                         int64_t pad_hi, unsigned flags, void *stream)
{
    int64_t total;
    if (int rc = check_common(in, out, dtype, ndim, shape, mode, &total)) return rc;
    if (axis < 0 || axis >= ndim) return set_error(B200_EINVAL, "axis out of range");
    if (size < 1 || size > (1 << 20)) return set_error(B200_EINVAL, "incorrect filter size");
    if (size / 2 + origin < 0 || size / 2 + origin >= size) return set_error(B200_EORIGIN, "invalid origin");
    if (pad_lo < 0 || pad_hi < 0 || ((pad_lo || pad_hi) && axis != 0))
        return set_error(B200_EINVAL, "halo pads are only defined for axis 0");
    if (total == 0) return B200_OK;
    MinMax1dArgs a;
    a.in = in; a.out = out; a.dtype = dtype;
    a.g = line_geometry(ndim, shape, axis);
    a.size = (int)size; a.is_max = is_max ? 1 : 0; a.mode = mode; a.cval = cval; a.origin = (int)origin;
    a.pad_lo = pad_lo; a.pad_hi = pad_hi; a.flags = flags; a.stream = (cudaStream_t)stream;
    if (!(flags & (B200_FLAG_EXACT | B200_FLAG_NO_TMA))) {
        int rc = launch_minmax1d_fast(a);
        if (rc >= 0) return rc;
    }
    if (flags & B200_FLAG_FORCE_FAST)
        return set_error(B200_EINVAL, "no tiled kernel is eligible for this call");
    return launch_minmax1d_generic(a);
}

int b200_minmax_filter_nd(const void *in, void *out, int dtype, int ndim, const int64_t *shape,
                          const uint8_t *footprint, const int64_t *fshape, const int64_t *origins, int is_max,
                          int mode, double cval, int64_t pad_lo, int64_t pad_hi, unsigned flags, void *stream)
{
    int64_t total;
    if (int rc = check_common(in, out, dtype, ndim, shape, mode, &total)) return rc;
    if (!footprint || !fshape || !origins) return set_error(B200_EINVAL, "no footprint provided");
    int64_t ftotal = 1;
    for (int d = 0; d < ndim; ++d) {
        if (fshape[d] < 1) return set_error(B200_EINVAL, "footprint must have non-zero extents");
        ftotal *= fshape[d];
        if (ftotal > (1 << 24)) return set_error(B200_EINVAL, "footprint too large");
        if (origins[d] < -(fshape[d] / 2) || origins[d] > (fshape[d] - 1) / 2)
            return set_error(B200_EORIGIN, "invalid origin");
    }
    if (pad_lo < 0 || pad_hi < 0) return set_error(B200_EINVAL, "negative halo pad");
    if (total == 0) return B200_OK;
    MinMaxNdArgs a;
    a.in = in; a.out = out; a.dtype = dtype; a.ndim = ndim;
    for (int d = 0; d < ndim; ++d) {
        a.shape[d] = shape[d]; a.fshape[d] = fshape[d]; a.origins[d] = origins[d];
    }
    a.footprint = footprint; a.is_max = is_max ? 1 : 0; a.mode = mode; a.cval = cval;
    a.pad_lo = pad_lo; a.pad_hi = pad_hi; a.flags = flags; a.stream = (cudaStream_t)stream;
    if (flags & B200_FLAG_FORCE_FAST)
        return set_error(B200_EINVAL, "no tiled kernel is eligible for this call");
    return launch_minmaxnd_generic(a);
}

int b200_combine(void *acc, const void *src, int dtype, int64_t n, int op, void *stream)
{
    if (dtype < B200_U8 || dtype > B200_F64) return set_error(B200_EDTYPE, "data type not supported");
    if (op < B200_EPI_STORE || op > B200_EPI_SQ_ACC_SQRT) return set_error(B200_EINVAL, "unknown combine op");
    if (n < 0) return set_error(B200_EINVAL, "negative length");
    if (n > 0 && (!acc || !src)) return set_error(B200_EINVAL, "null device pointer");
    return launch_combine(acc, src, dtype, n, op, (cudaStream_t)stream);
}

int b200_shutdown(void)
{
    std::lock_guard<std::mutex> lk(g_table_mu);
    for (auto &e : g_tables) cudaFree(e.dptr);
    g_tables.clear();
    return B200_OK;
}

const char *b200_last_kernel(void) { return tl_kernel; }
const char *b200_last_error(void) { return tl_error; }
const char *b200_version(void) { return "b200-ndfilters 0.1.0 (sm_100a)"; }

int64_t b200_launch_count(int reset)
{
    int64_t v = tl_launches;
    if (reset) tl_launches = 0;
    return v;
}

}  // extern "C"
