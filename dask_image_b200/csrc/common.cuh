// common.cuh -- shared device/host helpers of the B200 ndfilters kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <float.h>
#include <cmath>
#include "../../include/b200_ndfilters.h"

namespace b200 {

// ---------------------------------------------------------------------------
// error plumbing (thread-local; the C ABI is re-entrant)
// ---------------------------------------------------------------------------
int set_error(int code, const char *fmt, ...);
void note_kernel(const char *name);
void count_launch(int n = 1);
int check_cuda(cudaError_t e, const char *what);

#define B200_CUDA_TRY(expr)                                   \
    do {                                                      \
        int _rc = ::b200::check_cuda((expr), #expr);          \
        if (_rc) return _rc;                                  \
    } while (0)

// ---------------------------------------------------------------------------
// dtype helpers
// ---------------------------------------------------------------------------
__host__ __device__ inline int dtype_size(int dt)
{
    switch (dt) {
    case B200_U8: case B200_I8: return 1;
    case B200_U16: case B200_I16: return 2;
    case B200_U32: case B200_I32: case B200_F32: return 4;
    default: return 8;
    }
}

// Calls f.template operator()<T>() for the C type of `dt`.
template <typename F>
inline int dispatch_dtype(int dt, F &&f)
{
    switch (dt) {
    case B200_U8:  return f.template operator()<uint8_t>();
    case B200_I8:  return f.template operator()<int8_t>();
    case B200_U16: return f.template operator()<uint16_t>();
    case B200_I16: return f.template operator()<int16_t>();
    case B200_U32: return f.template operator()<uint32_t>();
    case B200_I32: return f.template operator()<int32_t>();
    case B200_U64: return f.template operator()<uint64_t>();
    case B200_I64: return f.template operator()<int64_t>();
    case B200_F32: return f.template operator()<float>();
    case B200_F64: return f.template operator()<double>();
    default: return set_error(B200_EDTYPE, "data type not supported");
    }
}

// ---------------------------------------------------------------------------
// boundary extension: position i (any integer) on an axis of length n ->
// source index in [0,n) or -1 ("use cval").  Closed forms of SURVEY.md A2;
// multiple reflections (radius >= n) are legal.
// ---------------------------------------------------------------------------
__host__ __device__ __forceinline__ int64_t ext_index(int64_t i, int64_t n, int mode)
{
    if (i >= 0 && i < n) return i;
    switch (mode) {
    case B200_MODE_NEAREST:
        return i < 0 ? 0 : n - 1;
    case B200_MODE_WRAP: {
        int64_t t = i % n;
        return t < 0 ? t + n : t;
    }
    case B200_MODE_REFLECT: {
        int64_t p = 2 * n;
        int64_t t = i % p;
        if (t < 0) t += p;
        return t < n ? t : p - 1 - t;
    }
    case B200_MODE_MIRROR: {
        if (n == 1) return 0;
        int64_t p = 2 * n - 2;
        int64_t t = i % p;
        if (t < 0) t += p;
        return t < n ? t : p - t;
    }
    default:
        return -1;
    }
}

// Slab form of the same map.  `pad_lo` / `pad_hi` real planes of the global
// volume are resident directly before / after the n planes this call produces
// output for, so the RESIDENT range is [-pad_lo, n + pad_hi).  Positions inside
// it are returned unchanged (possibly negative: they address neighbour planes);
// positions outside are folded by `mode` at the ends of the resident range --
// on an end slab that end is the global volume's boundary, on every other slab
// it is never reached (the pads cover the filter's reach).  pad_lo = pad_hi = 0
// is exactly ext_index().  Returns B200_USE_CVAL for "use cval".
#define B200_USE_CVAL INT64_MIN
__host__ __device__ __forceinline__ int64_t ext_index_padded(int64_t i, int64_t n, int mode,
                                                            int64_t pad_lo, int64_t pad_hi)
{
    if (i >= -pad_lo && i < n + pad_hi) return i;
    int64_t s = ext_index(i + pad_lo, n + pad_lo + pad_hi, mode);
    return s < 0 ? B200_USE_CVAL : s - pad_lo;
}

// ---------------------------------------------------------------------------
// double -> array dtype, reproducing what the reference's x86-64 build does
// with a plain C cast: integers truncate toward zero; 8/16-bit types and int32
// go through a 32-bit cvttsd2si (out of range / NaN -> 0x80000000, then the
// low bits are kept); uint32 / int64 through the 64-bit form; uint64 through
// gcc's two-range sequence.  SURVEY.md A4: u16 -1.5 -> 65535, 120000 -> 54464.
// ---------------------------------------------------------------------------
__host__ __device__ __forceinline__ int32_t x86_cvtt32(double v)
{
    // (one compare: a value in (-2^31 - 1, -2^31] truncates to INT32_MIN, the out-of-range result)
    return fabs(v) < 2147483648.0 ? (int32_t)v : INT32_MIN;
}
__host__ __device__ __forceinline__ int64_t x86_cvtt64(double v)
{
    return (v >= -9223372036854775808.0 && v < 9223372036854775808.0) ? (int64_t)v : INT64_MIN;
}

template <typename T> __host__ __device__ __forceinline__ T cast_out(double v);
template <> __host__ __device__ __forceinline__ uint8_t cast_out<uint8_t>(double v) { return (uint8_t)x86_cvtt32(v); }
template <> __host__ __device__ __forceinline__ int8_t cast_out<int8_t>(double v) { return (int8_t)x86_cvtt32(v); }
template <> __host__ __device__ __forceinline__ uint16_t cast_out<uint16_t>(double v) { return (uint16_t)x86_cvtt32(v); }
template <> __host__ __device__ __forceinline__ int16_t cast_out<int16_t>(double v) { return (int16_t)x86_cvtt32(v); }
template <> __host__ __device__ __forceinline__ int32_t cast_out<int32_t>(double v) { return x86_cvtt32(v); }
template <> __host__ __device__ __forceinline__ uint32_t cast_out<uint32_t>(double v) { return (uint32_t)x86_cvtt64(v); }
template <> __host__ __device__ __forceinline__ int64_t cast_out<int64_t>(double v) { return x86_cvtt64(v); }
template <> __host__ __device__ __forceinline__ uint64_t cast_out<uint64_t>(double v)
{
    if (v < 9223372036854775808.0) return (uint64_t)x86_cvtt64(v);
    return (uint64_t)x86_cvtt64(v - 9223372036854775808.0) ^ 0x8000000000000000ull;
}
template <> __host__ __device__ __forceinline__ float cast_out<float>(double v) { return (float)v; }
template <> __host__ __device__ __forceinline__ double cast_out<double>(double v) { return v; }

template <typename T> struct is_float_type { static constexpr bool value = false; };
template <> struct is_float_type<float> { static constexpr bool value = true; };
template <> struct is_float_type<double> { static constexpr bool value = true; };

// ---------------------------------------------------------------------------
// epilogues in the array dtype (numpy ufunc semantics: integer arithmetic
// wraps modulo 2^bits; sqrt of an integer array runs in the float type numpy's
// type resolution picks -- float16 for 8-bit, float32 for 16-bit, float64 for
// 32/64-bit -- and is cast back with truncation, casting='unsafe').
// ---------------------------------------------------------------------------
template <typename T> struct wrap_type { using type = T; };
template <> struct wrap_type<int8_t> { using type = uint8_t; };
template <> struct wrap_type<int16_t> { using type = uint16_t; };
template <> struct wrap_type<int32_t> { using type = uint32_t; };
template <> struct wrap_type<int64_t> { using type = uint64_t; };

template <typename T> __host__ __device__ __forceinline__ T t_add(T a, T b)
{
    using U = typename wrap_type<T>::type;
    return (T)(U)((U)a + (U)b);
}
template <typename T> __host__ __device__ __forceinline__ T t_mul(T a, T b)
{
    using U = typename wrap_type<T>::type;
    return (T)(U)((U)a * (U)b);
}
// (no contraction on the device: the intrinsics keep ptxas from fusing the pair)
#ifdef __CUDA_ARCH__
template <> __host__ __device__ __forceinline__ float t_add<float>(float a, float b) { return __fadd_rn(a, b); }
template <> __host__ __device__ __forceinline__ float t_mul<float>(float a, float b) { return __fmul_rn(a, b); }
template <> __host__ __device__ __forceinline__ double t_add<double>(double a, double b) { return __dadd_rn(a, b); }
template <> __host__ __device__ __forceinline__ double t_mul<double>(double a, double b) { return __dmul_rn(a, b); }
#else
template <> __host__ __device__ __forceinline__ float t_add<float>(float a, float b) { return a + b; }
template <> __host__ __device__ __forceinline__ float t_mul<float>(float a, float b) { return a * b; }
template <> __host__ __device__ __forceinline__ double t_add<double>(double a, double b) { return a + b; }
template <> __host__ __device__ __forceinline__ double t_mul<double>(double a, double b) { return a * b; }
#endif

template <typename T> __host__ __device__ __forceinline__ T t_sqrt(T a)
{
    if constexpr (sizeof(T) == 1) {
        // float16 loop: value -> half, sqrt evaluated in float, rounded to half
        __half h = __float2half_rn((float)a);
        __half s = __float2half_rn(sqrtf(__half2float(h)));
        return cast_out<T>((double)__half2float(s));
    } else if constexpr (sizeof(T) == 2) {
        return cast_out<T>((double)sqrtf((float)a));
    } else {
        return cast_out<T>(sqrt((double)a));
    }
}
template <> __host__ __device__ __forceinline__ float t_sqrt<float>(float a) { return sqrtf(a); }
template <> __host__ __device__ __forceinline__ double t_sqrt<double>(double a) { return sqrt(a); }

template <typename T> __host__ __device__ __forceinline__ T apply_epilogue(int epi, T prev, T r)
{
    switch (epi) {
    case B200_EPI_ACC: return t_add<T>(prev, r);
    case B200_EPI_SQ_STORE: return t_mul<T>(r, r);
    case B200_EPI_SQ_ACC: return t_add<T>(prev, t_mul<T>(r, r));
    case B200_EPI_SQ_ACC_SQRT: return t_sqrt<T>(t_add<T>(prev, t_mul<T>(r, r)));
    default: return r;
    }
}
__host__ __device__ inline bool epilogue_reads_out(int epi)
{
    return epi == B200_EPI_ACC || epi == B200_EPI_SQ_ACC || epi == B200_EPI_SQ_ACC_SQRT;
}

// ---------------------------------------------------------------------------
// geometry of a 1-D pass over a C-contiguous array: [outer][n][inner]
// ---------------------------------------------------------------------------
struct LineGeom {
    int64_t outer, n, inner;
};
inline LineGeom line_geometry(int ndim, const int64_t *shape, int axis)
{
    LineGeom g{1, shape[axis], 1};
    for (int d = 0; d < axis; ++d) g.outer *= shape[d];
    for (int d = axis + 1; d < ndim; ++d) g.inner *= shape[d];
    return g;
}

constexpr int kMaxTaps1D = 1024;  // taps carried in kernel parameters

// symmetry classification of the reference's compiled 1-D loop:
// +1 symmetric, -1 antisymmetric, 0 neither (or an even tap count)
inline int classify_symmetry(const double *w, int nw)
{
    if (!(nw & 1)) return 0;
    const int size1 = nw / 2;
    int sym = 1;
    for (int k = 1; k <= size1; ++k)
        if (std::fabs(w[size1 + k] - w[size1 - k]) > DBL_EPSILON) { sym = 0; break; }
    if (!sym) {
        sym = -1;
        for (int k = 1; k <= size1; ++k)
            if (std::fabs(w[size1 + k] + w[size1 - k]) > DBL_EPSILON) { sym = 0; break; }
    }
    return sym;
}

// host-side launch entry points implemented by the kernel translation units --
struct Corr1dArgs {
    const void *in;
    void *out;
    int dtype;
    LineGeom g;
    const double *w;  // host
    int nw;
    int mode;
    double cval;
    int origin;
    int64_t pad_lo, pad_hi;
    int epilogue;
    unsigned flags;
    cudaStream_t stream;
    // neighbour planes read in place (b200_correlate1d_halo); null = none
    const void *peer_lo = nullptr, *peer_hi = nullptr;
    int n_peer_lo = 0, n_peer_hi = 0;
};
int launch_corr1d_generic(const Corr1dArgs &a);
// axis-0 pass with the halo planes read where they live; -1 = geometry not eligible
int launch_corr1d_peer(const Corr1dArgs &a);
int corr1d_peer_eligible(int dtype, int64_t n, int64_t inner, int nw, int origin);
// returns B200_OK when it ran, -1 when the shape/alignment is not eligible
int launch_corr1d_tma(const Corr1dArgs &a);
// exact tiled path of the 8/16-bit integer dtypes (corr1d_int.cu); -1 = not eligible
int launch_corr1d_int(const Corr1dArgs &a);

// two consecutive float32 passes along the last two axes in one launch (corr2d_fused.cu):
// the array is viewed as [outer][ny][nx]
struct Corr2dArgs {
    const void *in;
    void *out;
    int dtype;
    int64_t outer, ny, nx;
    const double *w1;  // host: taps of the pass along axis ndim-2
    int nw1, origin1, mode1;
    const double *w2;  // host: taps of the pass along axis ndim-1
    int nw2, origin2, mode2;
    double cval;
    int epilogue;
    unsigned flags;
    cudaStream_t stream;
};
// -1 = no fused kernel for this pair of filters (the caller runs the two passes one by one)
int launch_corr2d_fused(const Corr2dArgs &a);
int corr2d_fused_eligible(int dtype, int64_t ny, int64_t nx, int nw1, int origin1, int nw2, int origin2);

struct Uniform1dArgs {
    const void *in;
    void *out;
    int dtype;
    LineGeom g;
    int size;
    int mode;
    double cval;
    int origin;
    int64_t pad_lo, pad_hi;
    unsigned flags;
    cudaStream_t stream;
};
int launch_uniform1d_generic(const Uniform1dArgs &a);
int launch_uniform1d_peer(const Uniform1dArgs &u, const void *peer_lo, int n_lo, const void *peer_hi, int n_hi);
int uniform1d_peer_eligible(int dtype, int64_t n, int64_t inner, int size, int origin, int mode, double cval);
int launch_uniform1d_fast(const Uniform1dArgs &a);
// can the exact float epilogue (tile_common.cuh: box_quotient_bits) serve this integer box filter?
bool uniform_box_ok(int dtype, int size, int mode, double cval);

// the box-filter passes along the last two axes of a u8 / u16 array in one launch (corr2d_fused.cu)
struct Uniform2dArgs {
    const void *in;
    void *out;
    int dtype;
    int64_t outer, ny, nx;
    int size1, origin1, mode1;   // axis ndim-2
    int size2, origin2, mode2;   // axis ndim-1
    double cval;
    unsigned flags;
    cudaStream_t stream;
};
// -1 = no fused kernel (the caller runs the two passes one by one)
int launch_uniform2d_fused(const Uniform2dArgs &a);
int uniform2d_fused_eligible(int dtype, int64_t ny, int64_t nx, int size1, int origin1, int mode1, int size2, int origin2,
                             int mode2, double cval);

struct CorrNdArgs {
    const void *in;
    void *out;
    int dtype;
    int ndim;
    int64_t shape[B200_MAXDIM];
    const double *w;  // host, C order, shape wshape
    int64_t wshape[B200_MAXDIM];
    int64_t origins[B200_MAXDIM];
    int mode;
    double cval;
    int64_t pad_lo, pad_hi;
    unsigned flags;
    cudaStream_t stream;
    // neighbour planes read in place (b200_correlate_nd_halo); null = none
    const void *peer_lo = nullptr, *peer_hi = nullptr;
    int n_peer_lo = 0, n_peer_hi = 0;
};
int launch_corrnd_generic(const CorrNdArgs &a);
// slab whose neighbour planes are read where they live; -1 = geometry not eligible
int launch_corrnd_peer(const CorrNdArgs &a);
int corrnd_peer_eligible(int dtype, int ndim, const int64_t *shape, const int64_t *wshape, const int64_t *origins);
int launch_corrnd_tiled(const CorrNdArgs &a);
// exact tiled path of the 8/16-bit integer dtypes, 2-D / 3-D (corrnd_int.cu); -1 = not eligible
int launch_corrnd_int(const CorrNdArgs &a);

int launch_combine(void *acc, const void *src, int dtype, int64_t n, int op, cudaStream_t stream);

struct MinMax1dArgs {
    const void *in;
    void *out;
    int dtype;
    LineGeom g;
    int size;
    int is_max;
    int mode;
    double cval;
    int origin;
    int64_t pad_lo, pad_hi;
    unsigned flags;
    cudaStream_t stream;
};
int launch_minmax1d_generic(const MinMax1dArgs &a);
int launch_minmax1d_fast(const MinMax1dArgs &a);   // -1 = not eligible

struct MinMaxNdArgs {
    const void *in;
    void *out;
    int dtype;
    int ndim;
    int64_t shape[B200_MAXDIM];
    const uint8_t *footprint;  // host, C order, shape fshape
    int64_t fshape[B200_MAXDIM];
    int64_t origins[B200_MAXDIM];
    int is_max;
    int mode;
    double cval;
    int64_t pad_lo, pad_hi;
    unsigned flags;
    cudaStream_t stream;
};
int launch_minmaxnd_generic(const MinMaxNdArgs &a);

int device_sm_count();

// Device copy of a small host table (compacted tap / offset tables of the kernels whose tables do not
// fit the parameter bank).  Tables are cached by content per device (a chunk-wise caller repeats the same
// filter on every chunk): a miss costs one cudaMalloc and one synchronous upload, a hit costs a lookup --
// no allocation, no copy and no host-device serialisation on the launch path.  At most 32 tables stay
// resident (least recently used first out; freed in b200_shutdown()).  nullptr = CUDA error (set).
const void *cached_device_table(const void *host, size_t bytes);

}  // namespace b200
