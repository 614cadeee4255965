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

template <typename TS, int KIND>
static int launch_int_typed(const IntPlan &pl, cudaStream_t stream)
{
    // (Two follow-ups of round 1 -- eight outputs per lane, and a certified fixed-point walk -- were
    // measured on the B200 in round 2 and were 7 % and 24 % SLOWER than these kernels
    // (profiles/r2_int_ab.log); they were removed.)
    if (pl.contig)
        corr1d_int_contig_kernel<TS, KIND><<<(unsigned)pl.blocks, kIntThreads, pl.smem_bytes, stream>>>(pl.p);
    else
        corr1d_int_strided_kernel<TS, KIND><<<(unsigned)pl.blocks, kIntThreads, pl.smem_bytes, stream>>>(pl.p);
    return check_cuda(cudaGetLastError(), "corr1d_int launch");
}

template <typename TS>
static int launch_int_kind(const IntPlan &pl, cudaStream_t stream)
{
    if (pl.kind > 0) return launch_int_typed<TS, 1>(pl, stream);
    if (pl.kind < 0) return launch_int_typed<TS, -1>(pl, stream);
    return launch_int_typed<TS, 0>(pl, stream);
}

// returns B200_OK when it ran, -1 when the call is not eligible
int launch_corr1d_int(const Corr1dArgs &a)
{
    IntPlan pl{};
    if (plan_corr1d_int(a, pl, classify_symmetry(a.w, a.nw)) != 0) return -1;
    if (pl.smem_bytes > 48 * 1024) return -1;
    int rc;
    switch (a.dtype) {
    case B200_U8: rc = launch_int_kind<uint8_t>(pl, a.stream); break;
    case B200_I8: rc = launch_int_kind<int8_t>(pl, a.stream); break;
    case B200_U16: rc = launch_int_kind<uint16_t>(pl, a.stream); break;
    default: rc = launch_int_kind<int16_t>(pl, a.stream); break;
    }
    if (rc) return rc;
    count_launch();
    note_kernel(pl.contig ? "corr1d_int_contig" : "corr1d_int_strided");
    return B200_OK;
}

}  // namespace b200
