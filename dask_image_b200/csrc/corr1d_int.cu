// corr1d_int.cu -- kernels and launcher of the exact tiled integer correlate1d
// (uint8 / int8 / uint16 / int16): gaussian_filter and its derivative,
// gradient-magnitude and laplace chains on integer images, sobel / prewitt /
// laplace, and every other 1-D weighted pass on those dtypes.  Bit-identical to
// scipy (see corr1d_int.cuh for the arithmetic); bounded by FP64 issue, not HBM:
// 2 * taps - 1 FP64 instructions per output (taps when the weights are
// symmetric or antisymmetric).
#include "corr1d_int.cuh"
#include <cstdlib>

namespace b200 {

template <typename TS, int KIND>
__global__ void __launch_bounds__(kIntThreads, 8)
corr1d_int_strided_kernel(const __grid_constant__ IntParams p)
{
    extern __shared__ __align__(16) unsigned char int_smem[];
    IntTap *wc = reinterpret_cast<IntTap *>(int_smem);
    int64_t *rows = reinterpret_cast<int64_t *>(int_smem + kIntTapBytes);
    TS *tile = reinterpret_cast<TS *>(int_smem + kIntTapBytes + kIntRowTableBytes);
    int_stage_taps(p, wc, threadIdx.x);
    int_strided_stage_rows(p, rows, threadIdx.x, blockIdx.x);
    __syncthreads();
    int_strided_stage<TS>(p, tile, rows, threadIdx.x, blockIdx.x);
    __syncthreads();
    int_strided_compute<TS, KIND>(p, tile, wc, threadIdx.x, blockIdx.x);
}

template <typename TS, int KIND>
__global__ void __launch_bounds__(kIntThreads, 8)
corr1d_int_contig_kernel(const __grid_constant__ IntParams p)
{
    extern __shared__ __align__(16) unsigned char int_smem[];
    IntTap *wc = reinterpret_cast<IntTap *>(int_smem);
    TS *tile = reinterpret_cast<TS *>(int_smem + kIntTapBytes);
    TS *otile = tile + kIntContigRows * p.pitch;
    int_stage_taps(p, wc, threadIdx.x);
    int_contig_stage<TS>(p, tile, threadIdx.x, blockIdx.x);
    __syncthreads();
    int_contig_compute<TS, KIND>(p, tile, otile, wc, threadIdx.x, blockIdx.x);
    __syncthreads();
    int_contig_writeout<TS>(p, otile, threadIdx.x, blockIdx.x);
}

#ifdef B200_INT_NQ8
// Experiment (build with NVCC_EXTRA=-DB200_INT_NQ8, select with B200_INT_NQ=8): eight outputs per lane
// and walk -- half the window loads, tap-table reads and loop overhead per output, twice the accumulators
// (6 CTAs per SM instead of 8).  Same phases, same arithmetic; tests/test_int_emul.py runs both widths.
template <typename TS, int KIND>
__global__ void __launch_bounds__(kIntThreads, 6)
corr1d_int_strided_kernel_nq8(const __grid_constant__ IntParams p)
{
    extern __shared__ __align__(16) unsigned char int_smem[];
    IntTap *wc = reinterpret_cast<IntTap *>(int_smem);
    int64_t *rows = reinterpret_cast<int64_t *>(int_smem + kIntTapBytes);
    TS *tile = reinterpret_cast<TS *>(int_smem + kIntTapBytes + kIntRowTableBytes);
    int_stage_taps(p, wc, threadIdx.x);
    int_strided_stage_rows(p, rows, threadIdx.x, blockIdx.x);
    __syncthreads();
    int_strided_stage<TS>(p, tile, rows, threadIdx.x, blockIdx.x);
    __syncthreads();
    int_strided_compute<TS, KIND, 8>(p, tile, wc, threadIdx.x, blockIdx.x);
}

template <typename TS, int KIND>
__global__ void __launch_bounds__(kIntThreads, 6)
corr1d_int_contig_kernel_nq8(const __grid_constant__ IntParams p)
{
    extern __shared__ __align__(16) unsigned char int_smem[];
    IntTap *wc = reinterpret_cast<IntTap *>(int_smem);
    TS *tile = reinterpret_cast<TS *>(int_smem + kIntTapBytes);
    TS *otile = tile + kIntContigRows * p.pitch;
    int_stage_taps(p, wc, threadIdx.x);
    int_contig_stage<TS>(p, tile, threadIdx.x, blockIdx.x);
    __syncthreads();
    int_contig_compute<TS, KIND, 8>(p, tile, otile, wc, threadIdx.x, blockIdx.x);
    __syncthreads();
    int_contig_writeout<TS>(p, otile, threadIdx.x, blockIdx.x);
}
#endif

#ifdef B200_INT_FIXED
// Experiment (build with NVCC_EXTRA=-DB200_INT_FIXED, select with B200_INT_FIXED=1): the certified
// fixed-point walk of corr1d_int.cuh -- one IMAD.WIDE per tap pair, the exact FP64 sequence only for the
// outputs whose integer sum lies within the proven error band of an integer.
template <typename TS, int KIND, int NQ = 4>
__global__ void __launch_bounds__(kIntThreads, NQ == 8 ? 6 : 8)
corr1d_int_strided_kernel_fx(const __grid_constant__ IntParams p, const __grid_constant__ IntFixedParams fx)
{
    extern __shared__ __align__(16) unsigned char int_smem[];
    IntTap *wc = reinterpret_cast<IntTap *>(int_smem);
    int64_t *rows = reinterpret_cast<int64_t *>(int_smem + kIntTapBytes);
    TS *tile = reinterpret_cast<TS *>(int_smem + kIntTapBytes + kIntRowTableBytes);
    int_stage_taps(p, wc, threadIdx.x);
    int_strided_stage_rows(p, rows, threadIdx.x, blockIdx.x);
    __syncthreads();
    int_strided_stage<TS>(p, tile, rows, threadIdx.x, blockIdx.x);
    __syncthreads();
    int_strided_compute_fixed<TS, KIND, NQ>(p, fx, tile, wc, threadIdx.x, blockIdx.x);
}

template <typename TS, int KIND, int NQ = 4>
__global__ void __launch_bounds__(kIntThreads, NQ == 8 ? 6 : 8)
corr1d_int_contig_kernel_fx(const __grid_constant__ IntParams p, const __grid_constant__ IntFixedParams fx)
{
    extern __shared__ __align__(16) unsigned char int_smem[];
    IntTap *wc = reinterpret_cast<IntTap *>(int_smem);
    TS *tile = reinterpret_cast<TS *>(int_smem + kIntTapBytes);
    TS *otile = tile + kIntContigRows * p.pitch;
    int_stage_taps(p, wc, threadIdx.x);
    int_contig_stage<TS>(p, tile, threadIdx.x, blockIdx.x);
    __syncthreads();
    int_contig_compute_fixed<TS, KIND, NQ>(p, fx, tile, otile, wc, threadIdx.x, blockIdx.x);
    __syncthreads();
    int_contig_writeout<TS>(p, otile, threadIdx.x, blockIdx.x);
}
#endif

template <typename TS, int KIND>
static int launch_int_typed(const IntPlan &pl, const IntFixedParams &fx, cudaStream_t stream)
{
#ifdef B200_INT_FIXED
    static const bool fixed = [] { const char *e = getenv("B200_INT_FIXED"); return e && atoi(e) == 1; }();
#ifdef B200_INT_NQ8
    static const bool fixed8 = [] { const char *e = getenv("B200_INT_NQ"); return e && atoi(e) == 8; }();
    if (fixed && fx.on && fixed8) {
        if (pl.contig)
            corr1d_int_contig_kernel_fx<TS, KIND, 8><<<(unsigned)pl.blocks, kIntThreads, pl.smem_bytes, stream>>>(pl.p, fx);
        else
            corr1d_int_strided_kernel_fx<TS, KIND, 8><<<(unsigned)pl.blocks, kIntThreads, pl.smem_bytes, stream>>>(pl.p, fx);
        return check_cuda(cudaGetLastError(), "corr1d_int launch");
    }
#endif
    if (fixed && fx.on) {
        if (pl.contig)
            corr1d_int_contig_kernel_fx<TS, KIND><<<(unsigned)pl.blocks, kIntThreads, pl.smem_bytes, stream>>>(pl.p, fx);
        else
            corr1d_int_strided_kernel_fx<TS, KIND><<<(unsigned)pl.blocks, kIntThreads, pl.smem_bytes, stream>>>(pl.p, fx);
        return check_cuda(cudaGetLastError(), "corr1d_int launch");
    }
#else
    (void)fx;
#endif
#ifdef B200_INT_NQ8
    static const bool nq8 = [] { const char *e = getenv("B200_INT_NQ"); return e && atoi(e) == 8; }();
    if (nq8) {
        if (pl.contig)
            corr1d_int_contig_kernel_nq8<TS, KIND><<<(unsigned)pl.blocks, kIntThreads, pl.smem_bytes, stream>>>(pl.p);
        else
            corr1d_int_strided_kernel_nq8<TS, KIND><<<(unsigned)pl.blocks, kIntThreads, pl.smem_bytes, stream>>>(pl.p);
        return check_cuda(cudaGetLastError(), "corr1d_int launch");
    }
#endif
    if (pl.contig)
        corr1d_int_contig_kernel<TS, KIND><<<(unsigned)pl.blocks, kIntThreads, pl.smem_bytes, stream>>>(pl.p);
    else
        corr1d_int_strided_kernel<TS, KIND><<<(unsigned)pl.blocks, kIntThreads, pl.smem_bytes, stream>>>(pl.p);
    return check_cuda(cudaGetLastError(), "corr1d_int launch");
}

template <typename TS>
static int launch_int_kind(const IntPlan &pl, const IntFixedParams &fx, cudaStream_t stream)
{
    if (pl.kind > 0) return launch_int_typed<TS, 1>(pl, fx, stream);
    if (pl.kind < 0) return launch_int_typed<TS, -1>(pl, fx, stream);
    return launch_int_typed<TS, 0>(pl, fx, stream);
}

// returns B200_OK when it ran, -1 when the call is not eligible
int launch_corr1d_int(const Corr1dArgs &a)
{
    IntPlan pl{};
    if (plan_corr1d_int(a, pl, classify_symmetry(a.w, a.nw)) != 0) return -1;
    if (pl.smem_bytes > 48 * 1024) return -1;
    IntFixedParams fx;
    fx.on = 0;
#ifdef B200_INT_FIXED
    plan_int_fixed(a, pl, fx);
#endif
    int rc;
    switch (a.dtype) {
    case B200_U8: rc = launch_int_kind<uint8_t>(pl, fx, a.stream); break;
    case B200_I8: rc = launch_int_kind<int8_t>(pl, fx, a.stream); break;
    case B200_U16: rc = launch_int_kind<uint16_t>(pl, fx, a.stream); break;
    default: rc = launch_int_kind<int16_t>(pl, fx, a.stream); break;
    }
    if (rc) return rc;
    count_launch();
    note_kernel(pl.contig ? "corr1d_int_contig" : "corr1d_int_strided");
    return B200_OK;
}

}  // namespace b200
