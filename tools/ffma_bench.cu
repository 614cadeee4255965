// Microbenchmark: FFMA vs FFMA2 (fma.rn.f32x2) issue/pipe throughput on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(256) k(float *out, float w0, float w1, int iters)
{
    float2 a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f);
    float2 x0 = make_float2(w0, w0 + 1.f), x1 = make_float2(w1, w1 * 2.f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                if (MODE == 0) {  // scalar FFMA x2
                    a[i].x = fmaf(a[i].x, w0, x0.x);
                    a[i].y = fmaf(a[i].y, w0, x0.y);
                } else if (MODE == 1) {  // FFMA2 with broadcast scalar weight
                    a[i] = __ffma2_rn(a[i], make_float2(w0, w0), x0);
                } else {  // FFMA2 with register pair weight
                    a[i] = __ffma2_rn(a[i], x1, x0);
                }
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i].x + a[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main()
{
    float *out;
    cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
    const int iters = 4000;
    for (int mode = 0; mode < 3; ++mode) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<148 * 8, 256>>>(out, 1.0001f, 0.5f, iters);
            if (mode == 1) k<1><<<148 * 8, 256>>>(out, 1.0001f, 0.5f, iters);
            if (mode == 2) k<2><<<148 * 8, 256>>>(out, 1.0001f, 0.5f, iters);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
        }
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        double fmas = 148.0 * 8 * 256 * (double)iters * 4 * 16 * 2;
        printf("mode %d: %.3f ms, %.2f T FMA lane-ops/s (%s)\n", mode, ms, fmas / ms / 1e9,
               mode == 0 ? "FFMA" : mode == 1 ? "FFMA2 bcast" : "FFMA2 pair");
    }
    return 0;
}
