// generic.cu -- shape-agnostic kernels of the ndfilters path (sm_100a).
//
// These cover every dtype, mode, origin, tap count and shape the reference
// accepts, accumulate in fp64 in exactly the order scipy's compiled loops use
// (no FMA contraction: __dmul_rn/__dadd_rn) and store with the reference's
// cast rules, so their results are bit-identical to scipy.ndimage for every
// dtype.  They are the path of the 32/64-bit integer dtypes (8/16-bit ones have exact
// tiled kernels: corr1d_int.cu, corrnd_int.cu), the B200_FLAG_EXACT path, and the
// fallback for shapes the TMA-tiled kernels (corr1d_strided.cu, corr1d_contig.cu,
// corrnd_tiled.cu) do not take.  Loads are coalesced along the contiguous
// axis and overlapping windows are served by L1/L2.
#include "common.cuh"
#include <cstring>
#include <vector>
#include <cmath>

namespace b200 {

// ===========================================================================
// correlate1d  (replaces _nd_image.correlate1d, scipy/ndimage/_filters.py:610)
// ===========================================================================
struct Corr1dGenericParams {
    const void *in;
    void *out;
    int64_t outer, n, inner, total;
    int64_t pad_lo, pad_hi;
    int nw, size1, size2, origin, mode, symmetric, epilogue;
    double cval;
    const double *wdev;  // used when nw > kMaxTaps1D
    double w[kMaxTaps1D];
};

template <typename T, typename IDX>
__global__ void __launch_bounds__(256)
corr1d_generic_kernel(const __grid_constant__ Corr1dGenericParams p)
{
    const T *__restrict__ in = (const T *)p.in;
    T *__restrict__ out = (T *)p.out;
    const double *w = p.wdev ? p.wdev : p.w;
    const double *fw = w + p.size1;
    const IDX total = (IDX)p.total, inner = (IDX)p.inner, n = (IDX)p.n;
    const IDX step = (IDX)gridDim.x * blockDim.x;
    for (IDX e = (IDX)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += step) {
        IDX c = e % inner;
        IDX t = e / inner;
        IDX i = t % n;
        IDX o = t / n;
        const T *line = in + ((int64_t)o * p.n * p.inner + (int64_t)c);
        const int64_t centre = (int64_t)i - p.origin;
        const bool interior = (centre - p.size1 >= -p.pad_lo) && (centre + p.size2 < p.n + p.pad_hi);
        auto X = [&](int64_t j) -> double {
            int64_t pos = centre + j;
            if (!interior) {
                pos = ext_index_padded(pos, p.n, p.mode, p.pad_lo, p.pad_hi);
                if (pos == B200_USE_CVAL) return p.cval;
            }
            return (double)line[pos * p.inner];
        };
        double acc;
        if (p.symmetric > 0) {
            acc = __dmul_rn(X(0), fw[0]);
            for (int j = -p.size1; j < 0; ++j)
                acc = __dadd_rn(acc, __dmul_rn(__dadd_rn(X(j), X(-j)), fw[j]));
        } else if (p.symmetric < 0) {
            acc = __dmul_rn(X(0), fw[0]);
            for (int j = -p.size1; j < 0; ++j)
                acc = __dadd_rn(acc, __dmul_rn(__dsub_rn(X(j), X(-j)), fw[j]));
        } else {
            acc = __dmul_rn(X(p.size2), fw[p.size2]);
            for (int j = -p.size1; j < p.size2; ++j)
                acc = __dadd_rn(acc, __dmul_rn(X(j), fw[j]));
        }
        T r = cast_out<T>(acc);
        if (p.epilogue != B200_EPI_STORE) {
            T prev = epilogue_reads_out(p.epilogue) ? out[e] : T(0);
            r = apply_epilogue<T>(p.epilogue, prev, r);
        }
        out[e] = r;
    }
}

static int grid_for(int64_t total, int threads)
{
    int64_t blocks = (total + threads - 1) / threads;
    int64_t cap = (int64_t)device_sm_count() * 32;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

int launch_corr1d_generic(const Corr1dArgs &a)
{
    Corr1dGenericParams p;
    p.in = a.in; p.out = a.out;
    p.outer = a.g.outer; p.n = a.g.n; p.inner = a.g.inner;
    p.total = a.g.outer * a.g.n * a.g.inner;
    p.pad_lo = a.pad_lo; p.pad_hi = a.pad_hi;
    p.nw = a.nw; p.size1 = a.nw / 2; p.size2 = a.nw - p.size1 - 1;
    p.origin = a.origin; p.mode = a.mode; p.epilogue = a.epilogue;
    p.symmetric = classify_symmetry(a.w, a.nw);
    p.cval = a.cval;
    p.wdev = nullptr;
    if (a.nw <= kMaxTaps1D) {
        for (int i = 0; i < a.nw; ++i) p.w[i] = a.w[i];
    } else {
        p.wdev = (const double *)cached_device_table(a.w, sizeof(double) * a.nw);
        if (!p.wdev) return B200_ECUDA;
    }
    const int threads = 256;
    const int blocks = grid_for(p.total, threads);
    const bool idx32 = p.total < (int64_t)0x7fffffff - (int64_t)blocks * threads;
    int rc = dispatch_dtype(a.dtype, [&]<typename T>() -> int {
        if (idx32)
            corr1d_generic_kernel<T, uint32_t><<<blocks, threads, 0, a.stream>>>(p);
        else
            corr1d_generic_kernel<T, int64_t><<<blocks, threads, 0, a.stream>>>(p);
        return check_cuda(cudaGetLastError(), "corr1d_generic launch");
    });
    if (rc) return rc;
    count_launch();
    note_kernel("corr1d_generic");
    return B200_OK;
}

// ===========================================================================
// uniform_filter1d (replaces _nd_image.uniform_filter1d, _filters.py:1542)
// out = (sum of the `size` window samples as doubles) / size.  The reference
// carries a running double sum along each line; window sums of integer data
// are exact in double (|sum| < 2^53), so the direct sum is bit-identical for
// integer dtypes; for floats it is the same sum up to double rounding.
// ===========================================================================
struct Uniform1dGenericParams {
    const void *in;
    void *out;
    int64_t outer, n, inner, total;
    int64_t pad_lo, pad_hi;
    int size, lo, mode;
    double cval;
};

template <typename T, typename IDX>
__global__ void __launch_bounds__(256)
uniform1d_generic_kernel(const __grid_constant__ Uniform1dGenericParams p)
{
    const T *__restrict__ in = (const T *)p.in;
    T *__restrict__ out = (T *)p.out;
    const IDX total = (IDX)p.total, inner = (IDX)p.inner, n = (IDX)p.n;
    const IDX step = (IDX)gridDim.x * blockDim.x;
    const double dsize = (double)p.size;
    for (IDX e = (IDX)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += step) {
        IDX c = e % inner;
        IDX t = e / inner;
        IDX i = t % n;
        IDX o = t / n;
        const T *line = in + ((int64_t)o * p.n * p.inner + (int64_t)c);
        const int64_t first = (int64_t)i - p.lo;
        const bool interior = (first >= -p.pad_lo) && (first + p.size - 1 < p.n + p.pad_hi);
        double sum = 0.0;
        if (interior) {
            for (int k = 0; k < p.size; ++k) sum = __dadd_rn(sum, (double)line[(first + k) * p.inner]);
        } else {
            for (int k = 0; k < p.size; ++k) {
                int64_t pos = ext_index_padded(first + k, p.n, p.mode, p.pad_lo, p.pad_hi);
                double v = (pos == B200_USE_CVAL) ? p.cval : (double)line[pos * p.inner];
                sum = __dadd_rn(sum, v);
            }
        }
        out[e] = cast_out<T>(__ddiv_rn(sum, dsize));
    }
}

int launch_uniform1d_generic(const Uniform1dArgs &a)
{
    Uniform1dGenericParams p;
    p.in = a.in; p.out = a.out;
    p.outer = a.g.outer; p.n = a.g.n; p.inner = a.g.inner;
    p.total = a.g.outer * a.g.n * a.g.inner;
    p.pad_lo = a.pad_lo; p.pad_hi = a.pad_hi;
    p.size = a.size; p.lo = a.size / 2 + a.origin; p.mode = a.mode; p.cval = a.cval;
    const int threads = 256;
    const int blocks = grid_for(p.total, threads);
    const bool idx32 = p.total < (int64_t)0x7fffffff - (int64_t)blocks * threads;
    int rc = dispatch_dtype(a.dtype, [&]<typename T>() -> int {
        if (idx32)
            uniform1d_generic_kernel<T, uint32_t><<<blocks, threads, 0, a.stream>>>(p);
        else
            uniform1d_generic_kernel<T, int64_t><<<blocks, threads, 0, a.stream>>>(p);
        return check_cuda(cudaGetLastError(), "uniform1d_generic launch");
    });
    if (rc) return rc;
    count_launch();
    note_kernel("uniform1d_generic");
    return B200_OK;
}

// ===========================================================================
// N-D correlate (replaces _nd_image.correlate, scipy/ndimage/_filters.py:1305)
// tmp = 0; for every tap with |w| > DBL_EPSILON in C order: tmp += w * x.
// ===========================================================================
struct CorrNdGenericParams {
    const void *in;
    void *out;
    int ndim;
    int64_t shape[B200_MAXDIM];
    int64_t stride[B200_MAXDIM];
    int64_t total;
    int64_t pad_lo, pad_hi;
    int lo[B200_MAXDIM], hi[B200_MAXDIM];
    int nnz;
    const double *tw;      // [nnz] weights
    const int32_t *tk;     // [nnz][ndim] relative tap coordinates
    const int64_t *tlin;   // [nnz] linear element offsets (interior fast path)
    int mode;
    double cval;
};

template <typename T, typename IDX>
__global__ void __launch_bounds__(256)
corrnd_generic_kernel(const __grid_constant__ CorrNdGenericParams p)
{
    const T *__restrict__ in = (const T *)p.in;
    T *__restrict__ out = (T *)p.out;
    const IDX total = (IDX)p.total;
    const IDX step = (IDX)gridDim.x * blockDim.x;
    const int nd = p.ndim;
    for (IDX e = (IDX)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += step) {
        int64_t idx[B200_MAXDIM];
        IDX r = e;
        bool interior = true;
#pragma unroll
        for (int d = B200_MAXDIM - 1; d >= 0; --d) {
            if (d < nd) {
                IDX s = (IDX)p.shape[d];
                IDX q = r / s;
                idx[d] = (int64_t)(r - q * s);
                r = q;
                int64_t plo = d == 0 ? p.pad_lo : 0, phi = d == 0 ? p.pad_hi : 0;
                interior = interior && (idx[d] - p.lo[d] >= -plo) && (idx[d] + p.hi[d] < p.shape[d] + phi);
            }
        }
        double tmp = 0.0;
        if (interior) {
            const T *base = in + (int64_t)e;
            for (int t = 0; t < p.nnz; ++t)
                tmp = __dadd_rn(tmp, __dmul_rn(p.tw[t], (double)base[p.tlin[t]]));
        } else {
            for (int t = 0; t < p.nnz; ++t) {
                int64_t off = 0;
                bool outside = false;
#pragma unroll
                for (int d = 0; d < B200_MAXDIM; ++d) {
                    if (d < nd) {
                        int64_t plo = d == 0 ? p.pad_lo : 0, phi = d == 0 ? p.pad_hi : 0;
                        int64_t s = ext_index_padded(idx[d] + p.tk[t * nd + d], p.shape[d], p.mode, plo, phi);
                        if (s == B200_USE_CVAL) outside = true;
                        else off += s * p.stride[d];
                    }
                }
                double v = outside ? p.cval : (double)in[off];
                tmp = __dadd_rn(tmp, __dmul_rn(p.tw[t], v));
            }
        }
        out[e] = cast_out<T>(tmp);
    }
}

int launch_corrnd_generic(const CorrNdArgs &a)
{
    CorrNdGenericParams p;
    p.in = a.in; p.out = a.out; p.ndim = a.ndim;
    p.total = 1;
    int64_t wtotal = 1;
    for (int d = 0; d < a.ndim; ++d) {
        p.shape[d] = a.shape[d];
        p.total *= a.shape[d];
        wtotal *= a.wshape[d];
        p.lo[d] = (int)(a.wshape[d] / 2 + a.origins[d]);
        p.hi[d] = (int)(a.wshape[d] - 1 - a.wshape[d] / 2 - a.origins[d]);
    }
    p.stride[a.ndim - 1] = 1;
    for (int d = a.ndim - 2; d >= 0; --d) p.stride[d] = p.stride[d + 1] * a.shape[d + 1];
    p.pad_lo = a.pad_lo; p.pad_hi = a.pad_hi; p.mode = a.mode; p.cval = a.cval;

    std::vector<double> tw;
    std::vector<int32_t> tk;
    std::vector<int64_t> tlin;
    for (int64_t f = 0; f < wtotal; ++f) {
        if (std::fabs(a.w[f]) > DBL_EPSILON) {
            int64_t r = f, lin = 0;
            int32_t k[B200_MAXDIM];
            for (int d = a.ndim - 1; d >= 0; --d) {
                k[d] = (int32_t)(r % a.wshape[d] - a.wshape[d] / 2 - a.origins[d]);
                r /= a.wshape[d];
                lin += (int64_t)k[d] * p.stride[d];
            }
            for (int d = 0; d < a.ndim; ++d) tk.push_back(k[d]);
            tw.push_back(a.w[f]);
            tlin.push_back(lin);
        }
    }
    p.nnz = (int)tw.size();
    const char *dev = nullptr;
    size_t b_tw = sizeof(double) * tw.size(), b_lin = sizeof(int64_t) * tlin.size(),
           b_tk = sizeof(int32_t) * tk.size();
    if (p.nnz > 0) {
        // one blob {weights, linear offsets, per-axis offsets}, cached on the device by content
        std::vector<unsigned char> blob(b_tw + b_lin + b_tk);
        memcpy(blob.data(), tw.data(), b_tw);
        memcpy(blob.data() + b_tw, tlin.data(), b_lin);
        memcpy(blob.data() + b_tw + b_lin, tk.data(), b_tk);
        dev = (const char *)cached_device_table(blob.data(), blob.size());
        if (!dev) return B200_ECUDA;
    }
    p.tw = (const double *)dev;
    p.tlin = (const int64_t *)(dev + b_tw);
    p.tk = (const int32_t *)(dev + b_tw + b_lin);

    const int threads = 256;
    const int blocks = grid_for(p.total, threads);
    const bool idx32 = p.total < (int64_t)0x7fffffff - (int64_t)blocks * threads;
    int rc = dispatch_dtype(a.dtype, [&]<typename T>() -> int {
        if (idx32)
            corrnd_generic_kernel<T, uint32_t><<<blocks, threads, 0, a.stream>>>(p);
        else
            corrnd_generic_kernel<T, int64_t><<<blocks, threads, 0, a.stream>>>(p);
        return check_cuda(cudaGetLastError(), "corrnd_generic launch");
    });
    if (rc) return rc;
    count_launch();
    note_kernel("corrnd_generic");
    return B200_OK;
}

// ===========================================================================
// minimum / maximum filters
//   1-D: replaces _nd_image.min_or_max_filter1d (scipy/ndimage/_filters.py:1660-1666,
//        1720-1726): values and cval go through double, the extreme of the
//        `size` window samples is cast back with the reference's cast rules.
//   N-D: replaces _nd_image.min_or_max_filter (scipy/ndimage/_filters.py:1771):
//        extreme over the true positions of a boolean footprint; here cval is
//        first cast to the ARRAY type (as the reference's C loop does).
// Min / max select, they do not round: results are bit-identical for every
// dtype (NaN ordering of the reference's ring-buffer scan is not reproduced).
// ===========================================================================
struct MinMax1dParams {
    const void *in;
    void *out;
    int64_t outer, n, inner, total;
    int64_t pad_lo, pad_hi;
    int size, lo, mode, is_max;
    double cval;
};

template <typename T, typename IDX>
__global__ void __launch_bounds__(256)
minmax1d_generic_kernel(const __grid_constant__ MinMax1dParams p)
{
    const T *__restrict__ in = (const T *)p.in;
    T *__restrict__ out = (T *)p.out;
    const IDX total = (IDX)p.total, inner = (IDX)p.inner, n = (IDX)p.n;
    const IDX step = (IDX)gridDim.x * blockDim.x;
    for (IDX e = (IDX)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += step) {
        IDX c = e % inner;
        IDX t = e / inner;
        IDX i = t % n;
        IDX o = t / n;
        const T *line = in + ((int64_t)o * p.n * p.inner + (int64_t)c);
        const int64_t first = (int64_t)i - p.lo;
        const bool interior = (first >= -p.pad_lo) && (first + p.size - 1 < p.n + p.pad_hi);
        double m = 0.0;
        for (int k = 0; k < p.size; ++k) {
            double v;
            if (interior) {
                v = (double)line[(first + k) * p.inner];
            } else {
                const int64_t pos = ext_index_padded(first + k, p.n, p.mode, p.pad_lo, p.pad_hi);
                v = (pos == B200_USE_CVAL) ? p.cval : (double)line[pos * p.inner];
            }
            if (k == 0) m = v;
            else if (p.is_max ? (v > m) : (v < m)) m = v;
        }
        out[e] = cast_out<T>(m);
    }
}

int launch_minmax1d_generic(const MinMax1dArgs &a)
{
    MinMax1dParams p;
    p.in = a.in; p.out = a.out;
    p.outer = a.g.outer; p.n = a.g.n; p.inner = a.g.inner;
    p.total = a.g.outer * a.g.n * a.g.inner;
    p.pad_lo = a.pad_lo; p.pad_hi = a.pad_hi;
    p.size = a.size; p.lo = a.size / 2 + a.origin; p.mode = a.mode; p.is_max = a.is_max; p.cval = a.cval;
    const int threads = 256;
    const int blocks = grid_for(p.total, threads);
    const bool idx32 = p.total < (int64_t)0x7fffffff - (int64_t)blocks * threads;
    int rc = dispatch_dtype(a.dtype, [&]<typename T>() -> int {
        if (idx32)
            minmax1d_generic_kernel<T, uint32_t><<<blocks, threads, 0, a.stream>>>(p);
        else
            minmax1d_generic_kernel<T, int64_t><<<blocks, threads, 0, a.stream>>>(p);
        return check_cuda(cudaGetLastError(), "minmax1d_generic launch");
    });
    if (rc) return rc;
    count_launch();
    note_kernel(a.is_max ? "max1d_generic" : "min1d_generic");
    return B200_OK;
}

struct MinMaxNdParams {
    const void *in;
    void *out;
    int ndim;
    int64_t shape[B200_MAXDIM];
    int64_t stride[B200_MAXDIM];
    int64_t total;
    int64_t pad_lo, pad_hi;
    int lo[B200_MAXDIM], hi[B200_MAXDIM];
    int nnz;
    const int32_t *tk;     // [nnz][ndim] relative tap coordinates
    const int64_t *tlin;   // [nnz] linear element offsets (interior fast path)
    int mode, is_max;
    double cval;
};

template <typename T, typename IDX>
__global__ void __launch_bounds__(256)
minmaxnd_generic_kernel(const __grid_constant__ MinMaxNdParams p)
{
    const T *__restrict__ in = (const T *)p.in;
    T *__restrict__ out = (T *)p.out;
    const IDX total = (IDX)p.total;
    const IDX step = (IDX)gridDim.x * blockDim.x;
    const int nd = p.ndim;
    const double cv = (double)cast_out<T>(p.cval);   // the reference casts cval to the array type first
    for (IDX e = (IDX)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += step) {
        int64_t idx[B200_MAXDIM];
        IDX r = e;
        bool interior = true;
#pragma unroll
        for (int d = B200_MAXDIM - 1; d >= 0; --d) {
            if (d < nd) {
                IDX s = (IDX)p.shape[d];
                IDX q = r / s;
                idx[d] = (int64_t)(r - q * s);
                r = q;
                int64_t plo = d == 0 ? p.pad_lo : 0, phi = d == 0 ? p.pad_hi : 0;
                interior = interior && (idx[d] - p.lo[d] >= -plo) && (idx[d] + p.hi[d] < p.shape[d] + phi);
            }
        }
        double m = 0.0;
        for (int t = 0; t < p.nnz; ++t) {
            double v;
            if (interior) {
                v = (double)in[(int64_t)e + p.tlin[t]];
            } else {
                int64_t off = 0;
                bool outside = false;
#pragma unroll
                for (int d = 0; d < B200_MAXDIM; ++d) {
                    if (d < nd) {
                        int64_t plo = d == 0 ? p.pad_lo : 0, phi = d == 0 ? p.pad_hi : 0;
                        int64_t s = ext_index_padded(idx[d] + p.tk[t * nd + d], p.shape[d], p.mode, plo, phi);
                        if (s == B200_USE_CVAL) outside = true;
                        else off += s * p.stride[d];
                    }
                }
                v = outside ? cv : (double)in[off];
            }
            if (t == 0) m = v;
            else if (p.is_max ? (v > m) : (v < m)) m = v;
        }
        out[e] = cast_out<T>(m);
    }
}

int launch_minmaxnd_generic(const MinMaxNdArgs &a)
{
    MinMaxNdParams p;
    p.in = a.in; p.out = a.out; p.ndim = a.ndim;
    p.total = 1;
    int64_t ftotal = 1;
    for (int d = 0; d < a.ndim; ++d) {
        p.shape[d] = a.shape[d];
        p.total *= a.shape[d];
        ftotal *= a.fshape[d];
        p.lo[d] = (int)(a.fshape[d] / 2 + a.origins[d]);
        p.hi[d] = (int)(a.fshape[d] - 1 - a.fshape[d] / 2 - a.origins[d]);
    }
    p.stride[a.ndim - 1] = 1;
    for (int d = a.ndim - 2; d >= 0; --d) p.stride[d] = p.stride[d + 1] * a.shape[d + 1];
    p.pad_lo = a.pad_lo; p.pad_hi = a.pad_hi; p.mode = a.mode; p.is_max = a.is_max; p.cval = a.cval;

    std::vector<int32_t> tk;
    std::vector<int64_t> tlin;
    for (int64_t f = 0; f < ftotal; ++f) {
        if (a.footprint[f]) {
            int64_t r = f, lin = 0;
            int32_t k[B200_MAXDIM];
            for (int d = a.ndim - 1; d >= 0; --d) {
                k[d] = (int32_t)(r % a.fshape[d] - a.fshape[d] / 2 - a.origins[d]);
                r /= a.fshape[d];
                lin += (int64_t)k[d] * p.stride[d];
            }
            for (int d = 0; d < a.ndim; ++d) tk.push_back(k[d]);
            tlin.push_back(lin);
        }
    }
    p.nnz = (int)tlin.size();
    if (p.nnz == 0) return set_error(B200_EINVAL, "All-zero footprint is not supported.");
    const size_t b_lin = sizeof(int64_t) * tlin.size(), b_tk = sizeof(int32_t) * tk.size();
    std::vector<unsigned char> blob(b_lin + b_tk);
    memcpy(blob.data(), tlin.data(), b_lin);
    memcpy(blob.data() + b_lin, tk.data(), b_tk);
    const char *dev = (const char *)cached_device_table(blob.data(), blob.size());
    if (!dev) return B200_ECUDA;
    p.tlin = (const int64_t *)dev;
    p.tk = (const int32_t *)(dev + b_lin);

    const int threads = 256;
    const int blocks = grid_for(p.total, threads);
    const bool idx32 = p.total < (int64_t)0x7fffffff - (int64_t)blocks * threads;
    int rc = dispatch_dtype(a.dtype, [&]<typename T>() -> int {
        if (idx32)
            minmaxnd_generic_kernel<T, uint32_t><<<blocks, threads, 0, a.stream>>>(p);
        else
            minmaxnd_generic_kernel<T, int64_t><<<blocks, threads, 0, a.stream>>>(p);
        return check_cuda(cudaGetLastError(), "minmaxnd_generic launch");
    });
    if (rc) return rc;
    count_launch();
    note_kernel(a.is_max ? "maxnd_generic" : "minnd_generic");
    return B200_OK;
}

// ===========================================================================
// elementwise combine (numpy ufuncs of scipy/ndimage/_filters.py:1025-1030,
// 1184-1193), in the array dtype
// ===========================================================================
template <typename T>
__global__ void __launch_bounds__(256)
combine_kernel(T *__restrict__ acc, const T *__restrict__ src, int64_t n, int op)
{
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += step) {
        T prev = epilogue_reads_out(op) ? acc[e] : T(0);
        acc[e] = apply_epilogue<T>(op, prev, src[e]);
    }
}

int launch_combine(void *acc, const void *src, int dtype, int64_t n, int op, cudaStream_t stream)
{
    if (n <= 0) return B200_OK;
    const int threads = 256;
    const int blocks = grid_for(n, threads);
    int rc = dispatch_dtype(dtype, [&]<typename T>() -> int {
        combine_kernel<T><<<blocks, threads, 0, stream>>>((T *)acc, (const T *)src, n, op);
        return check_cuda(cudaGetLastError(), "combine launch");
    });
    if (rc) return rc;
    count_launch();
    note_kernel("combine");
    return B200_OK;
}

}  // namespace b200
