// corr1d_contig.cuh -- tile staging shared by the contiguous-axis units
// (corr1d_contig.cu: weighted passes and box filter; corr1d_minmax.cu: min / max).
#pragma once
#include "corr1d_tma.cuh"

namespace b200 {

// ---------------------------------------------------------------------------
// Contiguous-axis tiles: the columns of a tile that fall outside [0, n) (TMA
// zero-fills them).  Edge CTAs fetch their images under the extension map while
// the tile is still in flight (kPre elements per thread into registers) and
// store them into the tile after the TMA has landed.
// ---------------------------------------------------------------------------
template <typename TS>
struct ContigEdge {
    int nlow, hstart, per_row, total;
    __device__ __forceinline__ void init(int64_t g0, int P, int64_t n, int TR)
    {
        nlow = g0 < 0 ? (int)(-g0 < P ? -g0 : P) : 0;
        hstart = g0 + P > n ? (int)((n - g0) > nlow ? (n - g0) : nlow) : P;
        per_row = nlow + (P - hstart);
        total = per_row * TR;
    }
    __device__ __forceinline__ int column(int k) const { return k < nlow ? k : hstart + (k - nlow); }
};

// LDG = stage with plain element loads instead of TMA (rows / bases that are not 16-byte multiples)
template <typename TS, int KPRE, bool LDG = false>
__device__ __forceinline__ void contig_stage_tile(const CUtensorMap &tmap, TS *tile, uint64_t *full_bar, const TS *in,
                                                  int64_t g0, int64_t row0, int64_t n, int64_t nrows, int P, int TR,
                                                  int mode, TS cval)
{
    const int tid = threadIdx.x;
    const int nthr = blockDim.x;
    if (!LDG && tid == 0) {
        mbar_init(full_bar, 1);
        fence_barrier_init();
        mbar_arrive_expect_tx(full_bar, (uint32_t)(TR * P * (int)sizeof(TS)));
        tma_load_2d(tile, &tmap, full_bar, (int)g0, (int)row0);
    }
    ContigEdge<TS> edge;
    edge.init(g0, P, n, TR);
    const bool patch = edge.total > 0 && !(mode == B200_MODE_CONSTANT && cval == TS(0));
    // source column of every patch column, computed once (the extension map
    // costs a 64-bit modulo): -1 = cval.  Rows are < 2^31 elements long.
    __shared__ int patch_src[256];
    if (patch) {
        for (int k = tid; k < edge.per_row; k += nthr) patch_src[k] = (int)ext_index(g0 + edge.column(k), n, mode);
        __syncthreads();
    }
    auto value = [&](int idx) -> TS {
        const int r = idx / edge.per_row, k = idx - r * edge.per_row;
        const int sc = patch_src[k];
        if (sc < 0) return cval;
        return (row0 + r < nrows) ? in[(row0 + r) * n + sc] : TS(0);
    };
    TS pre[KPRE];
    if (patch) {
#pragma unroll
        for (int i = 0; i < KPRE; ++i) {
            const int idx = tid + i * nthr;
            pre[i] = idx < edge.total ? value(idx) : TS(0);
        }
    }
    __syncthreads();
    if constexpr (LDG) {
        // the columns of the tile that lie inside the row, element by element; the rest is zero
        // (a warp walks along one tile row: coalesced, and no division per element)
        const int lane = tid & 31, wid = tid >> 5, nw = nthr >> 5;
        for (int r = wid; r < TR; r += nw) {
            const bool row_ok = row0 + r < nrows;
            const TS *src = in + (row0 + r) * n + g0;
            for (int c = lane; c < P; c += 32) {
                const int64_t col = g0 + c;
                const bool ok = row_ok && col >= 0 && col < n;
                if constexpr (sizeof(TS) == 4) {
                    cp_async_f32(tile + r * P + c, ok ? src + c : in, ok);
                } else {
                    tile[r * P + c] = ok ? src[c] : TS(0);
                }
            }
        }
        if constexpr (sizeof(TS) == 4) cp_async_wait_all();
    } else {
        if (tid < 32) mbar_wait(full_bar, 0);
    }
    __syncthreads();
    if (patch) {
#pragma unroll
        for (int i = 0; i < KPRE; ++i) {
            const int idx = tid + i * nthr;
            if (idx < edge.total) {
                const int r = idx / edge.per_row, k = idx - r * edge.per_row;
                tile[r * P + edge.column(k)] = pre[i];
            }
        }
        for (int idx = tid + KPRE * nthr; idx < edge.total; idx += nthr) {
            const int r = idx / edge.per_row, k = idx - r * edge.per_row;
            tile[r * P + edge.column(k)] = value(idx);
        }
        __syncthreads();
    }
}

}  // namespace b200
