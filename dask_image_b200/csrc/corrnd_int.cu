// corrnd_int.cu -- kernel and launcher of the exact tiled N-D correlate for
// uint8 / int8 / uint16 / int16 arrays (see corrnd_int.cuh).  Bounded by FP64
// issue: two FP64 instructions per tap and output.
#include "corrnd_int.cuh"

namespace b200 {

template <typename TS>
__global__ void __launch_bounds__(kIntThreads, 8)
corrnd_int_kernel(const __grid_constant__ IntNdParams p)
{
    extern __shared__ __align__(16) unsigned char int_nd_smem[];
    IntTap *wc = reinterpret_cast<IntTap *>(int_nd_smem);
    int64_t *table = reinterpret_cast<int64_t *>(int_nd_smem + int_nd_tap_bytes(p.ntaps));
    TS *tile = reinterpret_cast<TS *>(int_nd_smem + int_nd_tap_bytes(p.ntaps) + int_nd_table_bytes(p));
    TS *otile = tile + p.KZ * p.stage_lines * p.pitch;
    int_nd_stage_taps(p, wc, threadIdx.x);
    int_nd_stage_lines(p, table, threadIdx.x, blockIdx.x);
    __syncthreads();
    int_nd_stage<TS>(p, tile, table, threadIdx.x, blockIdx.x);
    __syncthreads();
    int_nd_compute<TS>(p, tile, otile, wc, threadIdx.x, blockIdx.x);
    __syncthreads();
    int_nd_writeout<TS>(p, otile, threadIdx.x, blockIdx.x);
}

template <typename TS>
static int launch_nd_typed(const IntNdPlan &pl, cudaStream_t stream)
{
    // (a constant, set on every launch: the attribute is per function AND per device)
    B200_CUDA_TRY(cudaFuncSetAttribute(corrnd_int_kernel<TS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)kIntNdMaxSmem));
    corrnd_int_kernel<TS><<<(unsigned)pl.blocks, kIntThreads, pl.smem_bytes, stream>>>(pl.p);
    return check_cuda(cudaGetLastError(), "corrnd_int launch");
}

// returns B200_OK when it ran, -1 when the call is not eligible
int launch_corrnd_int(const CorrNdArgs &a)
{
    IntNdPlan pl{};
    if (plan_corrnd_int(a, pl) != 0) return -1;
    // the tap table lives on the device, cached by content (no allocation or copy per launch)
    const IntTap *dev = (const IntTap *)cached_device_table(pl.taps.data(), sizeof(IntTap) * pl.taps.size());
    if (!dev) return B200_ECUDA;
    int rc;
    pl.p.taps = dev;
    switch (a.dtype) {
    case B200_U8: rc = launch_nd_typed<uint8_t>(pl, a.stream); break;
    case B200_I8: rc = launch_nd_typed<int8_t>(pl, a.stream); break;
    case B200_U16: rc = launch_nd_typed<uint16_t>(pl, a.stream); break;
    default: rc = launch_nd_typed<int16_t>(pl, a.stream); break;
    }
    if (rc) return rc;
    count_launch();
    note_kernel("corrnd_int");
    return B200_OK;
}

}  // namespace b200
