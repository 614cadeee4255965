// corr1d_strided.cu -- TMA-tiled 1-D pass along a NON-contiguous axis (sm_100a):
// the float hot path behind gaussian_filter and its derivative / gradient-magnitude /
// laplace variants for every axis but the last, the exact uint8 / uint16 box filter
// and min / max filter along such an axis, and the multi-GPU peer-halo pass.
//
// The array is viewed as [outer][n][inner]; one CTA stages a
// (tile_rows + taps - 1) x W box (rows along the filtered axis, W contiguous
// elements) with ONE cp.async.bulk.tensor.3d, then every thread owns a 16-byte
// column vector and R consecutive output rows, sliding a register window of R
// rows down the tile: per tap 1 LDS.128 + R*V FMAs.  corr1d_strided_peer_kernel
// is the same body for the axis-0 pass of a multi-GPU slab: the halo rows of the
// first / last tile come straight out of the neighbour GPUs' memory with extra
// TMA loads (PeerMaps).  Out-of-range rows are zero-filled by TMA; CTAs on an
// array edge fetch exactly those elements through the closed-form extension map
// while the tile is in flight, so interior CTAs carry no boundary logic at all.
#include "corr1d_tma.cuh"

namespace b200 {

// LDG = the tile is staged with plain element loads instead of TMA (rows or bases that are not
// 16-byte multiples) and stored element-wise: same tile, same arithmetic, no alignment needs.
template <typename T, typename TS, int W, int TY, int R, int REM, bool PEER, int OP = 0, bool LDG = false>
__device__ __forceinline__ void strided_tile_body(const CUtensorMap *tmap, const Corr1dTmaParams<T, TS> &p,
                                                  const PeerMaps *hm)
{
    constexpr int V = StridedV<T, TS>::V;
    constexpr bool kBox = sizeof(TS) != sizeof(T);   // exact integer box filter
    constexpr int TX = W / V;
    constexpr int NTHREADS = TX * TY;
    using VecT = Vec<T, V>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    TS *tile = aligned_smem<TS>(smem_raw);
    __shared__ __align__(8) uint64_t full_bar;

    const int tid = threadIdx.x;
    const int tx = tid % TX, ty = tid / TX;
    const int L = p.nw;
    const bool reads_out = epilogue_reads_out(p.epilogue);

    const int64_t t = blockIdx.x;
    const int tn = (int)(t % p.n_tiles);
    const int64_t q = t / p.n_tiles;
    const int tc = (int)(q % p.c_tiles);
    const int64_t o = q / p.c_tiles;
    const int64_t row0 = (int64_t)tn * p.tile_rows;  // first output row of this tile
    const int64_t col0 = (int64_t)tc * W;
    const int64_t g0 = row0 - p.lo;                  // array row held by smem row 0

    if (!LDG && tid == 0) {
        mbar_init(&full_bar, 1);
        fence_barrier_init();
        if constexpr (PEER) {
            // (the slab's tensor map has no pad rows in front: coordinates are slab rows)
            if (tn == 0 && hm->has_lo) {
                mbar_arrive_expect_tx(&full_bar, (uint32_t)(p.smem_rows * W * (int)sizeof(TS)));
                tma_load_3d(tile, &hm->lo, &full_bar, (int)col0, 0, 0);
                tma_load_3d(tile + hm->lo_rows * W, &hm->first, &full_bar, (int)col0, 0, 0);
            } else if (tn == p.n_tiles - 1 && hm->has_hi) {
                mbar_arrive_expect_tx(&full_bar, (uint32_t)((hm->k_last + hm->hi_rows) * W * (int)sizeof(TS)));
                tma_load_3d(tile, &hm->last, &full_bar, (int)col0, (int)g0, 0);
                tma_load_3d(tile + hm->k_last * W, &hm->hi, &full_bar, (int)col0, 0, 0);
            } else {
                mbar_arrive_expect_tx(&full_bar, (uint32_t)(p.smem_rows * W * (int)sizeof(TS)));
                tma_load_3d(tile, tmap, &full_bar, (int)col0, (int)g0, (int)o);
            }
        } else {
            mbar_arrive_expect_tx(&full_bar, (uint32_t)(p.smem_rows * W * (int)sizeof(TS)));
            tma_load_3d(tile, tmap, &full_bar, (int)col0, (int)(g0 + p.pad_lo), (int)o);
        }
    }

    // Rows of the tile that lie outside the resident range (TMA zero-fills
    // them): edge CTAs fetch their images under the extension map WHILE the
    // tile is in flight -- the first kPre elements per thread into registers --
    // and drop them into the tile once the TMA has landed, so the patch costs no
    // extra round trip to memory.  (Zeros are already right for constant/0.)
    constexpr int kPre = 8;
    const int64_t res_lo = -p.pad_lo, res_hi = p.n + p.pad_hi;        // resident rows
    const int nlo = (int)(g0 < res_lo ? (res_lo - g0 < p.smem_rows ? res_lo - g0 : p.smem_rows) : 0);
    const int rhi = (int)(g0 + p.smem_rows > res_hi ? (res_hi - g0 > nlo ? res_hi - g0 : nlo) : p.smem_rows);
    const int npatch_el = (nlo + (p.smem_rows - rhi)) * W;
    const bool patch = npatch_el > 0 && !(p.mode == B200_MODE_CONSTANT && p.cval == TS(0));
    auto patch_row = [&](int j) { return j < nlo ? j : rhi + (j - nlo); };
    // source row of every patch row, computed once (the extension map costs a
    // 64-bit modulo): -1 = cval
    __shared__ int patch_src[256];
    if (patch) {
        const int npr = nlo + (p.smem_rows - rhi);
        for (int k = tid; k < npr; k += NTHREADS) {
            const int64_t src = ext_index_padded(g0 + patch_row(k), p.n, p.mode, p.pad_lo, p.pad_hi);
            // resident rows relative to the first resident one, so the value fits an int
            patch_src[k] = src == B200_USE_CVAL ? -1 : (int)(src + p.pad_lo);
        }
        __syncthreads();
    }
    auto patch_value = [&](int idx) -> TS {
        const int j = idx / W, c = idx - j * W;
        const int sr = patch_src[j];
        if (sr < 0) return p.cval;
        if constexpr (PEER) {
            // The neighbours' rows are not in local memory.  A row that folds onto one of them lies
            // beyond the halo of a side that HAS a neighbour: it only feeds outputs past the slab's end,
            // which are never stored -- and the address next to the slab need not be mapped.
            if (sr < p.pad_lo || sr - p.pad_lo >= p.n) return TS(0);
        }
        return (col0 + c < p.inner) ? p.in[(o * p.n + (sr - p.pad_lo)) * p.inner + col0 + c] : TS(0);
    };
    TS pre[kPre];
    if (patch) {
#pragma unroll
        for (int i = 0; i < kPre; ++i) {
            const int idx = tid + i * NTHREADS;
            pre[i] = idx < npatch_el ? patch_value(idx) : TS(0);
        }
    }
    __syncthreads();
    if constexpr (LDG) {
        // resident rows of the tile, element by element (coalesced along the row); the rest is zero
        if constexpr (sizeof(TS) == 4) {
            // 4-byte asynchronous copies: the whole tile is in flight before anybody waits
            const TS *src = p.in + (o * p.n + g0) * p.inner + col0;
            for (int idx = tid; idx < p.smem_rows * W; idx += NTHREADS) {
                const int r = idx / W, c = idx - r * W;
                const bool ok = r >= nlo && r < rhi && col0 + c < p.inner;
                cp_async_f32(tile + idx, ok ? src + (int64_t)r * p.inner + c : p.in, ok);
            }
            cp_async_wait_all();
        } else {
            for (int idx = tid; idx < p.smem_rows * W; idx += NTHREADS) {
                const int r = idx / W, c = idx - r * W;
                TS v = TS(0);
                if (r >= nlo && r < rhi && col0 + c < p.inner) v = p.in[(o * p.n + (g0 + r)) * p.inner + col0 + c];
                tile[idx] = v;
            }
        }
    } else {
        // one warp polls the mbarrier; the others sleep at the CTA barrier
        if (tid < 32) mbar_wait(&full_bar, 0);
    }
    __syncthreads();
    if (patch) {
#pragma unroll
        for (int i = 0; i < kPre; ++i) {
            const int idx = tid + i * NTHREADS;
            if (idx < npatch_el) tile[patch_row(idx / W) * W + (idx % W)] = pre[i];
        }
        for (int idx = tid + kPre * NTHREADS; idx < npatch_el; idx += NTHREADS)
            tile[patch_row(idx / W) * W + (idx % W)] = patch_value(idx);
        __syncthreads();
    }

    const int Lfull = L - REM;
    const int64_t col = col0 + tx * V;
    TS *const out_col = p.out + (o * p.n * p.inner + col);
    if constexpr (kBox && OP == 0) {
        // Exact integer box filter: a SLIDING window down a run of consecutive rows.  A thread owns the
        // tile's rows [ty * run, (ty + 1) * run): it sums the first window once and then, per output
        // row, adds the row that enters and subtracts the row that leaves -- 2 loads and 2 packed adds
        // per output instead of `size`.  Window sums are integers below 2^23 held in floats, so every
        // order of additions and subtractions is exact: the result does not depend on where a run, a
        // tile or a slab starts (the partition-invariance a floating-point running sum would lose).
        const int run = p.nb * R;
        const int r0 = ty * run;
        const TS *sbase = tile + r0 * W + tx * V;
        typename TapAcc<T, V, OP>::type acc;
        acc.zero();
        for (int k = 0; k < L; ++k) acc.fma(load_block<T, TS, V>(sbase + k * W), T(1));
        for (int j = 0; j < run; ++j) {
            if (j > 0) {
                acc.fma(load_block<T, TS, V>(sbase + (L - 1 + j) * W), T(1));
                acc.fma(load_block<T, TS, V>(sbase + (j - 1) * W), T(-1));
            }
            const int64_t row = row0 + r0 + j;
            if (col < p.inner && row < p.n) {
                TS *dst = out_col + row * p.inner;
                if constexpr (LDG) {
#pragma unroll
                    for (int v = 0; v < V; ++v)
                        if (col + v < p.inner) dst[v] = (TS)box_quotient_bits<T>(acc.get(v), p.qinv, p.qc0);
                } else {
                    T sums[V];
#pragma unroll
                    for (int v = 0; v < V; ++v) sums[v] = acc.get(v);
                    store_box_quotients<T, TS, V>(dst, sums, p.qinv, p.qc0);
                }
            }
        }
        return;
    }
    for (int blk = 0; blk < p.nb; ++blk) {
        const int r0 = (blk * TY + ty) * R;
        const TS *sbase = tile + r0 * W + tx * V;
        VecT win[R];
        typename TapAcc<T, V, OP>::type acc[R];
#pragma unroll
        for (int j = 0; j < R; ++j) acc[j].zero();
#pragma unroll
        for (int j = 0; j < R - 1; ++j) win[j] = load_block<T, TS, V>(sbase + j * W);
        const TS *sk = sbase + (R - 1) * W;
        for (int k = 0; k < Lfull; k += R, sk += R * W) {
#pragma unroll
            for (int u = 0; u < R; ++u) {
                win[(u + R - 1) % R] = load_block<T, TS, V>(sk + u * W);
                const T wk = p.w[k + u];
#pragma unroll
                for (int j = 0; j < R; ++j) acc[j].fma(win[(u + j) % R], wk);
            }
        }
#pragma unroll
        for (int u = 0; u < REM; ++u) {
            win[(u + R - 1) % R] = load_block<T, TS, V>(sk + u * W);
            const T wk = p.w[Lfull + u];
#pragma unroll
            for (int j = 0; j < R; ++j) acc[j].fma(win[(u + j) % R], wk);
        }
        if (col < p.inner) {
#pragma unroll
            for (int j = 0; j < R; ++j) {
                const int64_t row = row0 + r0 + j;
                if (row < p.n) {
                    TS *dst = out_col + row * p.inner;
                    if constexpr (LDG) {
                        // unaligned rows: element stores, the row may end inside the vector
#pragma unroll
                        for (int v = 0; v < V; ++v) {
                            if (col + v < p.inner) {
                                if constexpr (kBox) {
                                    dst[v] = (TS)box_quotient_bits<T>(acc[j].get(v), p.qinv, p.qc0);
                                } else {
                                    const T prev = reads_out ? dst[v] : T(0);
                                    dst[v] = p.epilogue == B200_EPI_STORE
                                                 ? acc[j].get(v)
                                                 : apply_epilogue<T>(p.epilogue, prev, acc[j].get(v));
                                }
                            }
                        }
                    } else if constexpr (kBox) {
                        T sums[V];
#pragma unroll
                        for (int v = 0; v < V; ++v) sums[v] = acc[j].get(v);
                        store_box_quotients<T, TS, V>(dst, sums, p.qinv, p.qc0);
                    } else {
                        VecT res;
                        if (p.epilogue == B200_EPI_STORE) {
#pragma unroll
                            for (int v = 0; v < V; ++v) res.v[v] = acc[j].get(v);
                        } else {
                            VecT prev;
                            if (reads_out) prev = *reinterpret_cast<const VecT *>(dst);
#pragma unroll
                            for (int v = 0; v < V; ++v)
                                res.v[v] =
                                    apply_epilogue<T>(p.epilogue, reads_out ? prev.v[v] : T(0), acc[j].get(v));
                        }
                        *reinterpret_cast<VecT *>(dst) = res;
                    }
                }
            }
        }
    }
}


template <typename T, typename TS, int W, int TY, int R, int MINB, int REM>
__global__ void __launch_bounds__((W / StridedV<T, TS>::V) * TY, MINB)
corr1d_strided_tma_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ Corr1dTmaParams<T, TS> p)
{
    strided_tile_body<T, TS, W, TY, R, REM, false>(&tmap, p, nullptr);
}

// the same pass for rows / bases that are not 16-byte multiples: plain loads and stores
template <typename T, typename TS, int W, int TY, int R, int MINB, int REM>
__global__ void __launch_bounds__((W / StridedV<T, TS>::V) * TY, MINB)
corr1d_strided_ldg_kernel(const __grid_constant__ Corr1dTmaParams<T, TS> p)
{
    strided_tile_body<T, TS, W, TY, R, REM, false, 0, true>(nullptr, p, nullptr);
}

// minimum (OP = 1) / maximum (OP = 2) filter along a non-contiguous axis: the
// same tile, window walk and stores, the taps select instead of accumulating
template <typename TS, int W, int TY, int R, int MINB, int REM, int OP>
__global__ void __launch_bounds__((W / StridedV<float, TS>::V) * TY, MINB)
minmax_strided_tma_kernel(const __grid_constant__ CUtensorMap tmap,
                          const __grid_constant__ Corr1dTmaParams<float, TS> p)
{
    strided_tile_body<float, TS, W, TY, R, REM, false, OP>(&tmap, p, nullptr);
}

template <typename T, typename TS, int W, int TY, int R, int MINB, int REM>
__global__ void __launch_bounds__((W / StridedV<T, TS>::V) * TY, MINB)
corr1d_strided_peer_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ PeerMaps hm,
                           const __grid_constant__ Corr1dTmaParams<T, TS> p)
{
    strided_tile_body<T, TS, W, TY, R, REM, true>(&tmap, p, &hm);
}

// ---------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------
template <typename T, typename TS, int R, int MINB, int REM, int TY = 8, int OP = 0, bool LDG = false>
static int launch_strided_cfg(const Corr1dArgs &a, const BoxInfo &box, int nb)
{
    constexpr int V = StridedV<T, TS>::V;
    constexpr int W = 32 * V;   // one warp spans a tile row (512 bytes of floats)
    constexpr int ROWS_PER_BLK = TY * R;
    const int es = (int)sizeof(TS);
    const LineGeom &g = a.g;
    const int L = a.nw;
    const int64_t n_tot = a.pad_lo + g.n + a.pad_hi;
    const char *base = (const char *)a.in - a.pad_lo * g.inner * es;

    const int c_tiles = (int)((g.inner + W - 1) / W);
    while (nb > 1 && (nb * ROWS_PER_BLK + L - 1 > 256)) --nb;
    if (nb * ROWS_PER_BLK + L - 1 > 256) return -1;
    // do not make tiles taller than the axis
    while (nb > 1 && (int64_t)(nb - 1) * ROWS_PER_BLK >= g.n) --nb;
    const int tile_rows = nb * ROWS_PER_BLK;
    const int smem_rows = tile_rows + L - 1;
    const size_t stage_bytes = (size_t)smem_rows * W * es;
    if (stage_bytes + 256 > kSmemPerSm - 1024) return -1;
    const int64_t n_tiles = (g.n + tile_rows - 1) / tile_rows;
    const int64_t num_tiles = n_tiles * c_tiles * g.outer;
    if (num_tiles >= (1ll << 31)) return -1;

    CUtensorMap tmap;
    if constexpr (!LDG) {
        uint64_t dims[3] = {(uint64_t)g.inner, (uint64_t)n_tot, (uint64_t)g.outer};
        uint64_t strides[2] = {(uint64_t)g.inner * es, (uint64_t)g.inner * es * (uint64_t)n_tot};
        uint32_t bx[3] = {(uint32_t)W, (uint32_t)smem_rows, 1u};
        if (encode_tensor_map(&tmap, dtype_code<TS>(), (void *)base, 3, dims, strides, bx) != 0) return -1;
    }

    Corr1dTmaParams<T, TS> p;
    p.in = (const TS *)a.in; p.out = (TS *)a.out;
    p.outer = g.outer; p.n = g.n; p.inner = g.inner;
    p.pad_lo = a.pad_lo; p.pad_hi = a.pad_hi;
    p.nw = L; p.lo = a.nw / 2 + a.origin; p.mode = a.mode; p.epilogue = a.epilogue;
    p.tile_rows = tile_rows; p.smem_rows = smem_rows; p.nb = nb; p.pitch = W; p.shift = 0; p.ncg = 0;
    p.wide_store = 0;
    p.n_tiles = (int)n_tiles; p.c_tiles = c_tiles;
    p.cval = (TS)a.cval;
    p.qinv = (T)box.qinv; p.qc0 = (T)box.qc0;
    for (int i = 0; i < L; ++i) p.w[i] = (T)a.w[i];

    const size_t smem = stage_bytes + 128;
    if constexpr (LDG) {
        static_assert(OP == 0, "plain-load staging is built for the weighted pass and the box filter");
        auto kern = corr1d_strided_ldg_kernel<T, TS, W, TY, R, MINB, REM>;
        B200_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemOptIn));
        kern<<<(unsigned)num_tiles, (W / V) * TY, smem, a.stream>>>(p);
    } else if constexpr (OP == 0) {
        auto kern = corr1d_strided_tma_kernel<T, TS, W, TY, R, MINB, REM>;
        B200_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemOptIn));
        kern<<<(unsigned)num_tiles, (W / V) * TY, smem, a.stream>>>(tmap, p);
    } else {
        auto kern = minmax_strided_tma_kernel<TS, W, TY, R, MINB, REM, OP>;
        B200_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemOptIn));
        kern<<<(unsigned)num_tiles, (W / V) * TY, smem, a.stream>>>(tmap, p);
    }
    B200_CUDA_TRY(cudaGetLastError());
    count_launch();
    note_kernel(OP == 0 ? kernel_name<T, TS>(true, LDG) : (OP == 1 ? "min1d_tma_strided" : "max1d_tma_strided"));
    return B200_OK;
}

template <typename T, typename TS, int R, int OP = 0>
static int launch_strided_r(const Corr1dArgs &a, const BoxInfo &box)
{
    constexpr int MINB = 3;
    constexpr int V = StridedV<T, TS>::V;
    const int es = (int)sizeof(TS);
    const int row_bytes = 32 * V * es;
    const LineGeom &g = a.g;
    const int L = a.nw;
    if (L > kMaxTapsTma || 8 * R + L - 1 > 256) return -1;
    const int64_t n_tot = a.pad_lo + g.n + a.pad_hi;
    if (g.inner >= (1ll << 31) || n_tot >= (1ll << 31) || g.outer >= (1ll << 31)) return -1;
    if ((a.pad_lo || a.pad_hi) && g.outer != 1) return -1;   // pads exist only for axis 0
    const char *base = (const char *)a.in - a.pad_lo * g.inner * es;
    // TMA needs 16-byte rows and bases; anything else takes the same tiles with plain loads / stores
    // (the choice depends on the row length and the pointers only, never on the plane count)
    const bool unaligned = (g.inner * es) % 16 != 0 || ((uintptr_t)base & 15) || ((uintptr_t)a.out & 15);
    if (unaligned) {
        if constexpr (R == 4 && OP == 0) {
            int nbl = ((int)((kSmemPerSm - 3072) / MINB) / row_bytes - (L - 1)) / (8 * R);
            nbl = nbl < 1 ? 1 : nbl;
            switch (L % R) {
            case 0: return launch_strided_cfg<T, TS, R, MINB, 0, 8, 0, true>(a, box, nbl);
            case 1: return launch_strided_cfg<T, TS, R, MINB, 1, 8, 0, true>(a, box, nbl);
            case 2: return launch_strided_cfg<T, TS, R, MINB, 2, 8, 0, true>(a, box, nbl);
            default: return launch_strided_cfg<T, TS, R, MINB, 3, 8, 0, true>(a, box, nbl);
            }
        } else {
            return -1;
        }
    }

    // nb = row blocks (of 8*R rows) per CTA tile: the tallest tile that still
    // leaves MINB CTAs resident per SM (B200 sweeps in profiles/: more, smaller
    // CTAs beat a deeper per-CTA pipeline for this streaming kernel)
    // Long filters are FP32-issue bound rather than HBM bound: a fourth resident
    // CTA hides more of each CTA's load phase (B200 sweep, profiles/).
    // CTA shape: 8 warps (8*R rows per row block) or, for R = 4, 4 warps with
    // twice the CTAs resident (env B200_TMA_TY / B200_TMA_CTAS for sweeps)
    // B200 sweep (tools/sweep_strided2.py): short filters are HBM bound and like
    // 3 big CTAs; from ~24 taps on the pass is FP32-issue bound and 6-8 small
    // CTAs overlap their load and FMA phases better.
    const int ty = (R == 4 && OP == 0) ? env_int("B200_TMA_TY", (L >= 24 && sizeof(T) == sizeof(TS)) ? 4 : 8) : 8;
    int nb = env_int("B200_TMA_NB", 0);
    if (nb <= 0) {
        int ctas = env_int("B200_TMA_CTAS", 0);
        if (ctas <= 0) ctas = ty == 4 ? (L < 32 ? 8 : 6) : MINB;
        const int per_cta = (int)((kSmemPerSm - 3072) / ctas);
        const int max_rows = per_cta / row_bytes;
        nb = (max_rows - (L - 1)) / (ty * R);
        nb = nb < 1 ? 1 : nb;
    }
    if constexpr (R == 1) {
        return launch_strided_cfg<T, TS, R, MINB, 0>(a, box, nb);
    } else if constexpr (R == 2) {
        return (L & 1) ? launch_strided_cfg<T, TS, R, MINB, 1>(a, box, nb)
                       : launch_strided_cfg<T, TS, R, MINB, 0>(a, box, nb);
    } else if constexpr (OP != 0) {
        switch (L % R) {
        case 0: return launch_strided_cfg<T, TS, R, MINB, 0, 8, OP>(a, box, nb);
        case 1: return launch_strided_cfg<T, TS, R, MINB, 1, 8, OP>(a, box, nb);
        case 2: return launch_strided_cfg<T, TS, R, MINB, 2, 8, OP>(a, box, nb);
        default: return launch_strided_cfg<T, TS, R, MINB, 3, 8, OP>(a, box, nb);
        }
    } else {
        if (ty == 4) {
            switch (L % R) {
            case 0: return launch_strided_cfg<T, TS, R, MINB, 0, 4>(a, box, nb);
            case 1: return launch_strided_cfg<T, TS, R, MINB, 1, 4>(a, box, nb);
            case 2: return launch_strided_cfg<T, TS, R, MINB, 2, 4>(a, box, nb);
            default: return launch_strided_cfg<T, TS, R, MINB, 3, 4>(a, box, nb);
            }
        }
        switch (L % R) {
        case 0: return launch_strided_cfg<T, TS, R, MINB, 0>(a, box, nb);
        case 1: return launch_strided_cfg<T, TS, R, MINB, 1>(a, box, nb);
        case 2: return launch_strided_cfg<T, TS, R, MINB, 2>(a, box, nb);
        default: return launch_strided_cfg<T, TS, R, MINB, 3>(a, box, nb);
        }
    }
}

// Tile height (in row blocks of 32) for the peer-halo pass: the tallest tile
// <= nb_max blocks for which neighbour rows enter only the first and the last
// tile of the axis.  0 = no such geometry.
static int peer_row_blocks(int64_t n, int L, int lo, int hi, int nb_max)
{
    for (int nb = nb_max; nb >= 1; --nb) {
        const int tile_rows = nb * 32;
        if (tile_rows + L - 1 > 256) continue;
        const int64_t n_tiles = (n + tile_rows - 1) / tile_rows;
        const int64_t last_rows = n - (n_tiles - 1) * tile_rows;
        if (n_tiles >= 2 && tile_rows >= lo && last_rows >= hi && n >= tile_rows + hi) return nb;
    }
    return 0;
}

// Axis-0 pass of a slab whose neighbour planes are read in place (PeerMaps).
// float32 / float64, R = 4, 8-warp CTAs.  Returns -1 when the geometry does not
// fit (the caller then exchanges the planes and uses the padded form).
template <typename T, typename TS, int REM>
static int launch_peer_cfg(const Corr1dArgs &a, const BoxInfo &box, int nb)
{
    constexpr int R = 4, MINB = 3, TY = 8;
    constexpr int V = StridedV<T, TS>::V;
    constexpr int W = 32 * V;
    constexpr int ROWS_PER_BLK = TY * R;
    const int es = (int)sizeof(TS);
    const LineGeom &g = a.g;
    const int L = a.nw;
    const int lo = L / 2 + a.origin, hi = L - 1 - lo;
    const bool has_lo = a.peer_lo != nullptr && lo > 0, has_hi = a.peer_hi != nullptr && hi > 0;
    if ((has_lo && a.n_peer_lo != lo) || (has_hi && a.n_peer_hi != hi)) return -1;

    const int c_tiles = (int)((g.inner + W - 1) / W);
    nb = peer_row_blocks(g.n, L, lo, hi, nb);
    if (nb < 1) return -1;
    const int tile_rows = nb * ROWS_PER_BLK;
    const int smem_rows = tile_rows + L - 1;
    const size_t stage_bytes = (size_t)smem_rows * W * es;
    if (stage_bytes + 256 > kSmemPerSm - 1024) return -1;
    const int64_t n_tiles = (g.n + tile_rows - 1) / tile_rows;
    const int64_t last_rows = g.n - (n_tiles - 1) * tile_rows;
    const int k_last = (int)(last_rows + lo);
    const int64_t num_tiles = n_tiles * c_tiles;
    if (num_tiles >= (1ll << 31)) return -1;

    CUtensorMap tmap;
    PeerMaps hm;
    memset(&hm, 0, sizeof(hm));
    uint64_t dims[3] = {(uint64_t)g.inner, (uint64_t)g.n, 1u};
    uint64_t strides[2] = {(uint64_t)g.inner * es, (uint64_t)g.inner * es * (uint64_t)g.n};
    uint32_t bx[3] = {(uint32_t)W, (uint32_t)smem_rows, 1u};
    if (encode_tensor_map(&tmap, a.dtype, (void *)a.in, 3, dims, strides, bx) != 0) return -1;
    hm.has_lo = has_lo; hm.has_hi = has_hi; hm.lo_rows = lo; hm.hi_rows = hi; hm.k_last = k_last;
    if (has_lo) {
        uint32_t b1[3] = {(uint32_t)W, (uint32_t)(smem_rows - lo), 1u};
        if (encode_tensor_map(&hm.first, a.dtype, (void *)a.in, 3, dims, strides, b1) != 0) return -1;
        uint64_t d2[3] = {(uint64_t)g.inner, (uint64_t)lo, 1u};
        uint64_t s2[2] = {(uint64_t)g.inner * es, (uint64_t)g.inner * es * (uint64_t)lo};
        uint32_t b2[3] = {(uint32_t)W, (uint32_t)lo, 1u};
        if (encode_tensor_map(&hm.lo, a.dtype, (void *)a.peer_lo, 3, d2, s2, b2) != 0) return -1;
    }
    if (has_hi) {
        uint32_t b1[3] = {(uint32_t)W, (uint32_t)k_last, 1u};
        if (encode_tensor_map(&hm.last, a.dtype, (void *)a.in, 3, dims, strides, b1) != 0) return -1;
        uint64_t d2[3] = {(uint64_t)g.inner, (uint64_t)hi, 1u};
        uint64_t s2[2] = {(uint64_t)g.inner * es, (uint64_t)g.inner * es * (uint64_t)hi};
        uint32_t b2[3] = {(uint32_t)W, (uint32_t)hi, 1u};
        if (encode_tensor_map(&hm.hi, a.dtype, (void *)a.peer_hi, 3, d2, s2, b2) != 0) return -1;
    }

    Corr1dTmaParams<T, TS> p;
    p.in = (const TS *)a.in; p.out = (TS *)a.out;
    p.outer = 1; p.n = g.n; p.inner = g.inner;
    p.pad_lo = has_lo ? lo : 0; p.pad_hi = has_hi ? hi : 0;   // rows the neighbours supply: never patched
    p.nw = L; p.lo = lo; p.mode = a.mode; p.epilogue = a.epilogue;
    p.tile_rows = tile_rows; p.smem_rows = smem_rows; p.nb = nb; p.pitch = W; p.shift = 0; p.ncg = 0;
    p.wide_store = 0;
    p.n_tiles = (int)n_tiles; p.c_tiles = c_tiles;
    p.cval = (TS)a.cval;
    p.qinv = (T)box.qinv; p.qc0 = (T)box.qc0;
    for (int i = 0; i < L; ++i) p.w[i] = (T)a.w[i];

    auto kern = corr1d_strided_peer_kernel<T, TS, W, TY, R, MINB, REM>;
    const size_t smem = stage_bytes + 128;
    B200_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemOptIn));
    kern<<<(unsigned)num_tiles, (W / V) * TY, smem, a.stream>>>(tmap, hm, p);
    B200_CUDA_TRY(cudaGetLastError());
    count_launch();
    note_kernel(sizeof(TS) != sizeof(T) ? (sizeof(TS) == 2 ? "uniform1d_tma_strided_peer_u16"
                                                           : "uniform1d_tma_strided_peer_u8")
                                        : (sizeof(T) == 4 ? "corr1d_tma_strided_peer_f32"
                                                          : "corr1d_tma_strided_peer_f64"));
    return B200_OK;
}

// geometry test of launch_peer_cfg, without launching (all ranks of a slab
// decomposition must agree on the path: they evaluate it for every slab size)
int corr1d_peer_eligible(int dtype, int64_t n, int64_t inner, int nw, int origin)
{
    if (dtype != B200_F32 && dtype != B200_F64 && dtype != B200_U16 && dtype != B200_U8) return 0;
    constexpr int R = 4, MINB = 3, ROWS_PER_BLK = 8 * R;
    const int es = dtype_size(dtype);
    const int L = nw;
    if (L < 1 || origin < -(L / 2) || origin > (L - 1) / 2) return 0;
    const int lo = L / 2 + origin, hi = L - 1 - lo;
    if (inner == 1 || (inner * es) % 16 != 0 || inner >= (1ll << 31) || n >= (1ll << 31)) return 0;
    if (L > kMaxTapsTma || 8 * R + L - 1 > 256) return 0;
    int nb = ((int)((kSmemPerSm - 3072) / MINB) / (es == 1 ? 256 : 512) - (L - 1)) / ROWS_PER_BLK;
    nb = peer_row_blocks(n, L, lo, hi, nb < 1 ? 1 : nb);
    if (nb < 1) return 0;
    const int tile_rows = nb * ROWS_PER_BLK;
    const int64_t n_tiles = (n + tile_rows - 1) / tile_rows;
    const int wtile = 32 * (es >= 4 ? 16 / es : 8);   // tile width in elements
    const int64_t c_tiles = (inner + wtile - 1) / wtile;
    return n_tiles * c_tiles < (1ll << 31) ? 1 : 0;
}

template <typename T, typename TS>
static int launch_peer_typed(const Corr1dArgs &a, const BoxInfo &box)
{
    constexpr int R = 4, MINB = 3;
    const int es = (int)sizeof(TS);
    const LineGeom &g = a.g;
    const int L = a.nw;
    if (g.outer != 1 || g.inner == 1) return -1;
    if (L > kMaxTapsTma || 8 * R + L - 1 > 256) return -1;
    if ((g.inner * es) % 16 != 0) return -1;
    if (g.inner >= (1ll << 31) || g.n >= (1ll << 31)) return -1;
    if (((uintptr_t)a.in & 15) || ((uintptr_t)a.out & 15) || ((uintptr_t)a.peer_lo & 15) || ((uintptr_t)a.peer_hi & 15))
        return -1;
    const int per_cta = (int)((kSmemPerSm - 3072) / MINB);
    int nb = (per_cta / (32 * StridedV<T, TS>::V * es) - (L - 1)) / (8 * R);
    nb = nb < 1 ? 1 : nb;
    switch (L % R) {
    case 0: return launch_peer_cfg<T, TS, 0>(a, box, nb);
    case 1: return launch_peer_cfg<T, TS, 1>(a, box, nb);
    case 2: return launch_peer_cfg<T, TS, 2>(a, box, nb);
    default: return launch_peer_cfg<T, TS, 3>(a, box, nb);
    }
}

int launch_corr1d_peer(const Corr1dArgs &a)
{
    const BoxInfo none;
    if (a.dtype == B200_F32) return launch_peer_typed<float, float>(a, none);
    if (a.dtype == B200_F64) return launch_peer_typed<double, double>(a, none);
    return -1;
}

// R = output rows per thread; a tile is a multiple of 8 * R rows.  Thin calls
// (the edge planes of a slab, filtered after the halo exchange) take fewer rows
// per thread so that tiles are not mostly padding.  Every output still sums its
// taps in the same order, so the choice does not change a single bit.
template <typename T, typename TS>
static int launch_strided(const Corr1dArgs &a, const BoxInfo &box)
{
    const int es = (int)sizeof(TS);
    const char *base = (const char *)a.in - a.pad_lo * a.g.inner * es;
    const bool unaligned = (a.g.inner * es) % 16 != 0 || ((uintptr_t)base & 15) || ((uintptr_t)a.out & 15);
    if constexpr (sizeof(T) == sizeof(TS)) {
        if (!unaligned && a.g.n <= 8) return launch_strided_r<T, TS, 1>(a, box);
        if (!unaligned && a.g.n <= 16) return launch_strided_r<T, TS, 2>(a, box);
    }
    return launch_strided_r<T, TS, 4>(a, box);
}

// ---------------------------------------------------------------------------
// entry points (corr1d_tma.cuh)
// ---------------------------------------------------------------------------
int tma_strided_corr(const Corr1dArgs &a, const BoxInfo &box)
{
    switch (a.dtype) {
    case B200_F32: return launch_strided<float, float>(a, box);
    case B200_F64: return launch_strided<double, double>(a, box);
    case B200_U16: return launch_strided<float, uint16_t>(a, box);
    case B200_U8: return launch_strided<float, uint8_t>(a, box);
    default: return -1;
    }
}

int tma_strided_minmax(const Corr1dArgs &a, const BoxInfo &box, bool is_max)
{
    switch (a.dtype) {
    case B200_F32: return is_max ? launch_strided_r<float, float, 4, 2>(a, box) : launch_strided_r<float, float, 4, 1>(a, box);
    case B200_U16: return is_max ? launch_strided_r<float, uint16_t, 4, 2>(a, box) : launch_strided_r<float, uint16_t, 4, 1>(a, box);
    case B200_U8: return is_max ? launch_strided_r<float, uint8_t, 4, 2>(a, box) : launch_strided_r<float, uint8_t, 4, 1>(a, box);
    default: return -1;
    }
}

int tma_peer_box(const Corr1dArgs &a, const BoxInfo &box)
{
    if (a.dtype == B200_U16) return launch_peer_typed<float, uint16_t>(a, box);
    if (a.dtype == B200_U8) return launch_peer_typed<float, uint8_t>(a, box);
    return -1;
}

}  // namespace b200
