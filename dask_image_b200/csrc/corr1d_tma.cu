// corr1d_tma.cu -- TMA-staged shared-memory 1-D passes (sm_100a): the float hot
// path behind gaussian_filter and its derivative / gradient-magnitude / laplace
// variants (one launch per 1-D pass of scipy's chain,
// scipy/ndimage/_filters.py:852-858), the float box filter, and the exact
// uint8 / uint16 box filter.
//
//  * strided_tile_body / corr1d_strided_tma_kernel -- the filtered axis is NOT
//    the contiguous one.  The array is viewed as [outer][n][inner]; one CTA
//    stages a (tile_rows + taps - 1) x W box (rows along the filtered axis, W
//    contiguous elements) with ONE cp.async.bulk.tensor.3d, then every thread
//    owns a 16-byte column vector and R consecutive output rows, sliding a
//    register window of R rows down the tile: per tap 1 LDS.128 + R*V FMAs.
//    corr1d_strided_peer_kernel is the same body for the axis-0 pass of a
//    multi-GPU slab: the halo rows of the first / last tile come straight out
//    of the neighbour GPUs' memory with extra TMA loads (PeerMaps).
//
//  * corr1d_contig_static_kernel -- the filtered axis IS the contiguous one,
//    float32 (or uint8/uint16 storage) and at most 48 taps: the tap count is a
//    template parameter and the tap loop is straight-line packed FFMA2.
//    corr1d_contig_tma_kernel is the run-time-loop form (float64, longer
//    filters).  View [rows][n]; a CTA stages a TR x P box with one
//    cp.async.bulk.tensor.2d whose pitch P is an odd multiple of 16 bytes, so
//    the lanes of a quarter warp hit 8 distinct 16-byte bank groups.
//
// Out-of-range rows/columns are zero-filled by TMA; CTAs on an array edge
// fetch exactly those elements through the closed-form extension map (reflect /
// mirror / nearest / wrap / constant) while the tile is in flight, so interior
// CTAs carry no boundary logic at all.  Accumulation is in the array's own
// float type (fp32 for f32, fp64 for f64) with FMA; B200_FLAG_EXACT selects
// generic.cu.  The integer box filter sums in fp32 (exact below 2^23) and
// divides with one proven-exact fused rounding (box_quotient_bits).
#include "tile_common.cuh"
#include <mutex>
#include <cstdlib>
#include <cstring>
#include <cmath>

namespace b200 {

// ---------------------------------------------------------------------------
// host: tensor-map encoder through the runtime's driver entry point lookup
// ---------------------------------------------------------------------------
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                        const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                        CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                        CUtensorMapFloatOOBfill);

static PFN_tmapEncodeTiled tmap_encoder()
{
    static PFN_tmapEncodeTiled fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (PFN_tmapEncodeTiled)p;
    });
    return fn;
}

int encode_tensor_map(CUtensorMap *map, int dtype, void *base, int rank, const uint64_t *dims,
                      const uint64_t *strides_bytes, const uint32_t *box)
{
    PFN_tmapEncodeTiled enc = tmap_encoder();
    if (!enc) return -1;
    CUtensorMapDataType dt;
    switch (dtype) {
    case B200_U8: case B200_I8: dt = CU_TENSOR_MAP_DATA_TYPE_UINT8; break;
    case B200_U16: case B200_I16: dt = CU_TENSOR_MAP_DATA_TYPE_UINT16; break;
    case B200_U32: dt = CU_TENSOR_MAP_DATA_TYPE_UINT32; break;
    case B200_I32: dt = CU_TENSOR_MAP_DATA_TYPE_INT32; break;
    case B200_U64: dt = CU_TENSOR_MAP_DATA_TYPE_UINT64; break;
    case B200_I64: dt = CU_TENSOR_MAP_DATA_TYPE_INT64; break;
    case B200_F32: dt = CU_TENSOR_MAP_DATA_TYPE_FLOAT32; break;
    case B200_F64: dt = CU_TENSOR_MAP_DATA_TYPE_FLOAT64; break;
    default: return -1;
    }
    cuuint64_t gdim[3];
    cuuint64_t gstr[2];
    cuuint32_t bdim[3], estr[3];
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bdim[i] = box[i];
        estr[i] = 1;
        if (i + 1 < rank) gstr[i] = strides_bytes[i];
    }
    CUresult r = enc(map, dt, (cuuint32_t)rank, base, gdim, gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : -1;
}

// ---------------------------------------------------------------------------
constexpr int kMaxTapsTma = 208;

// T = arithmetic type (float / double), TS = storage type of the array: TS == T
// for the float passes; TS = uint8_t / uint16_t with T = float for the exact
// integer box filter (see the integer epilogue below).
template <typename T, typename TS = T>
struct Corr1dTmaParams {
    const TS *in;   // plane/row 0 of the array (edge patch-up loads)
    TS *out;
    int64_t outer, n, inner;
    int64_t pad_lo, pad_hi;
    int nw, lo, mode, epilogue;
    int tile_rows;   // strided: outputs per CTA tile along the filtered axis
    int smem_rows;   // strided: tile_rows + nw - 1
    int nb;          // strided: row blocks per thread
    int pitch;       // contig: smem row pitch in elements (V * odd)
    int shift;       // contig: leading taps to skip (box start aligned down to 16 bytes)
    int ncg;         // contig: 4*R-column groups per tile
    int wide_store;  // contig: rows and base are 32-byte aligned -> 256-bit stores
    int n_tiles, c_tiles;
    TS cval;
    T qinv, qc0;     // integer box filter: 1/size and the (0.5 - size/2) offset
    T w[kMaxTapsTma];
};

// ===========================================================================
// strided-axis pass: one tile per CTA, 3 CTAs resident per SM
// ===========================================================================
// REM = taps % R is a template parameter so that the tap tail is straight-line
// code; nothing in the tap loops is divergent, which lets ptxas keep the
// weights on the uniform datapath (LDCU -> FFMA2 with a UR operand).
// Peer halos (multi-GPU axis-0 pass, sharded.py): the planes below / above the
// slab are NOT copied next to it first -- the kernel reads them where they
// live, in the neighbour GPU's memory, with its own TMA loads over NVLink, so
// the "exchange" is part of the tile pipeline of the pass itself: no NCCL
// kernel competing for SMs, no staging planes, no separate edge launches.
// The first tile of the axis takes its top `lo` rows from `lo` (the previous
// rank's last planes) and the rest from a shorter local box; the last tile takes
// the `hi` rows behind the slab from `hi`.  Every other tile is unchanged.
struct PeerMaps {
    CUtensorMap first;   // local slab, box rows = smem_rows - lo (first tile, below its peer rows)
    CUtensorMap last;    // local slab, box rows = k_last         (last tile, above which peer rows follow)
    CUtensorMap lo;      // the lo planes that precede the slab (in a peer's memory)
    CUtensorMap hi;      // the hi planes that follow it
    int has_lo, has_hi;
    int lo_rows, hi_rows, k_last;
};

template <typename T, typename TS, int W, int TY, int R, int REM, bool PEER, int OP = 0>
__device__ __forceinline__ void strided_tile_body(const CUtensorMap &tmap, const Corr1dTmaParams<T, TS> &p,
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

    if (tid == 0) {
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
                tma_load_3d(tile, &tmap, &full_bar, (int)col0, (int)g0, (int)o);
            }
        } else {
            mbar_arrive_expect_tx(&full_bar, (uint32_t)(p.smem_rows * W * (int)sizeof(TS)));
            tma_load_3d(tile, &tmap, &full_bar, (int)col0, (int)(g0 + p.pad_lo), (int)o);
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
    // one warp polls the mbarrier; the others sleep at the CTA barrier
    if (tid < 32) mbar_wait(&full_bar, 0);
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
                    if constexpr (kBox) {
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
    strided_tile_body<T, TS, W, TY, R, REM, false>(tmap, p, nullptr);
}

// minimum (OP = 1) / maximum (OP = 2) filter along a non-contiguous axis: the
// same tile, window walk and stores, the taps select instead of accumulating
template <typename TS, int W, int TY, int R, int MINB, int REM, int OP>
__global__ void __launch_bounds__((W / StridedV<float, TS>::V) * TY, MINB)
minmax_strided_tma_kernel(const __grid_constant__ CUtensorMap tmap,
                          const __grid_constant__ Corr1dTmaParams<float, TS> p)
{
    strided_tile_body<float, TS, W, TY, R, REM, false, OP>(tmap, p, nullptr);
}

template <typename T, typename TS, int W, int TY, int R, int MINB, int REM>
__global__ void __launch_bounds__((W / StridedV<T, TS>::V) * TY, MINB)
corr1d_strided_peer_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ PeerMaps hm,
                           const __grid_constant__ Corr1dTmaParams<T, TS> p)
{
    strided_tile_body<T, TS, W, TY, R, REM, true>(tmap, p, &hm);
}

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

template <typename TS, int KPRE>
__device__ __forceinline__ void contig_stage_tile(const CUtensorMap &tmap, TS *tile, uint64_t *full_bar, const TS *in,
                                                  int64_t g0, int64_t row0, int64_t n, int64_t nrows, int P, int TR,
                                                  int mode, TS cval)
{
    const int tid = threadIdx.x;
    const int nthr = blockDim.x;
    if (tid == 0) {
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
    if (tid < 32) mbar_wait(full_bar, 0);
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

// ===========================================================================
// contiguous-axis pass: one tile per CTA
// ===========================================================================
// SH = leading taps to skip (box start aligned down to 16 bytes), REM = taps % V
template <typename T, typename TS, int SH, int REM>
__global__ void __launch_bounds__(256, 4)
corr1d_contig_tma_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ Corr1dTmaParams<T, TS> p)
{
    constexpr int V = StorageV<T, TS>::V;
    constexpr bool kBox = sizeof(TS) != sizeof(T);   // exact integer box filter
    constexpr int R = 2 * V;        // outputs per thread
    constexpr int PC = 4 * R;       // patch columns (4 lanes along x)
    using VecT = Vec<T, V>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    TS *tile = aligned_smem<TS>(smem_raw);
    __shared__ __align__(8) uint64_t full_bar;

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int P = p.pitch;
    const int TR = p.tile_rows;      // rows per tile (multiple of 8)
    const int SX = p.ncg * PC;       // tile columns
    const int nrg = TR / 8;
    const int L = p.nw;              // taps incl. the SH alignment taps
    const int64_t nrows = p.outer;
    const bool reads_out = epilogue_reads_out(p.epilogue);

    const int64_t t = blockIdx.x;
    const int tcx = (int)(t % p.c_tiles);      // tile index along the filtered (contiguous) axis
    const int64_t trow = t / p.c_tiles;
    const int64_t x0t = (int64_t)tcx * SX;     // first output column of this tile
    const int64_t row0 = trow * TR;
    const int64_t g0 = x0t - p.lo;             // array column held by smem column 0 (16-byte aligned)

    contig_stage_tile<TS, 8>(tmap, tile, &full_bar, p.in, g0, row0, p.n, nrows, P, TR, p.mode, p.cval);

    // Patches: a warp keeps ONE column group (ncg is 1, 2 or 4) and walks the row
    // groups with a fixed stride -- no div/mod, pointers advance by constants.
    // Every warp runs the same trip count (surplus iterations recompute row
    // group 0 and skip the store) so the tap loops stay free of divergence and
    // the weights stay on the uniform datapath.
    const int cg_shift = p.ncg == 4 ? 2 : (p.ncg == 2 ? 1 : 0);
    const int cg = warp & (p.ncg - 1);
    const int rg0 = warp >> cg_shift;
    const int rg_step = 8 >> cg_shift;
    const int patch_iters = (nrg + rg_step - 1) / rg_step;
    const int xl = cg * PC + (lane & 3) * R;
    const int r_in = lane >> 2;
    const int64_t x = x0t + xl;
    const TS *const tile_thr = tile + r_in * P + xl;
    TS *const out_thr = p.out + ((row0 + r_in) * p.n + x);
    for (int qi = 0; qi < patch_iters; ++qi) {
        const int rg_raw = rg0 + qi * rg_step;
        const bool patch_live = rg_raw < nrg;
        const int rg = patch_live ? rg_raw : 0;
        const int64_t row = row0 + rg * 8 + r_in;
        const TS *srow = tile_thr + rg * 8 * P;
        ContigAcc<T, V> acc;
        acc.zero();
        row_taps<T, TS, V, SH, REM>(acc, srow, p.w, L);

        if (patch_live && row < nrows) {
            TS *dst0 = out_thr + (int64_t)(rg * 8) * p.n;
            if constexpr (kBox) {
                // rows are whole multiples of 16 bytes, so x < n implies x + R <= n
                if (x < p.n) {
                    T sums[R];
#pragma unroll
                    for (int j = 0; j < R; ++j) sums[j] = acc.get(j);
                    store_box_quotients<T, TS, R>(dst0, sums, p.qinv, p.qc0);
                }
            } else if (p.wide_store && x + R <= p.n) {
                // whole 32-byte sector(s) in one store
                T res[R];
                if (p.epilogue == B200_EPI_STORE) {
#pragma unroll
                    for (int j = 0; j < R; ++j) res[j] = acc.get(j);
                } else {
                    T prev[R];
                    if (reads_out) ld_global_256(dst0, prev);
#pragma unroll
                    for (int j = 0; j < R; ++j)
                        res[j] = apply_epilogue<T>(p.epilogue, reads_out ? prev[j] : T(0), acc.get(j));
                }
                st_global_256(dst0, res);
            } else {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int64_t xx = x + h * V;
                    if (xx < p.n) {
                        TS *dst = dst0 + h * V;
                        VecT res;
                        if (p.epilogue == B200_EPI_STORE) {
#pragma unroll
                            for (int v = 0; v < V; ++v) res.v[v] = acc.get(h * V + v);
                        } else {
                            VecT prev;
                            if (reads_out) prev = *reinterpret_cast<const VecT *>(dst);
#pragma unroll
                            for (int v = 0; v < V; ++v)
                                res.v[v] =
                                    apply_epilogue<T>(p.epilogue, reads_out ? prev.v[v] : T(0), acc.get(h * V + v));
                        }
                        *reinterpret_cast<VecT *>(dst) = res;
                    }
                }
            }
        }
    }
}

// ===========================================================================
// contiguous-axis pass, float32, tap count fixed at compile time
// ===========================================================================
// NT = taps including the alignment zeros in front (the TMA box starts on a
// 16-byte boundary).  The whole tap loop is straight-line code: the thread's
// 16 + NT input floats of the row come in as LDS.128, every tap is a run of
// packed FFMA2 (8 for an even data offset, 9 for an odd one: the odd set pairs
// outputs (-1,0)...(15,16) so operand pairs stay register aligned) with its
// weight read through the uniform datapath at a static parameter offset.  No
// loop-carried register rotation, no per-group branches: ~14 instructions per
// output instead of ~38 for the run-time loop.  A thread owns 16 outputs, a
// warp 64 x 8; the lanes of a quarter warp sit on 8 different rows, which the
// odd 16-byte pitch spreads over all bank groups.  TS = uint16_t / uint8_t is
// the exact integer box filter (taps 0/1, quotient epilogue).
template <typename TS, int NT>
__global__ void __launch_bounds__(256, 3)
corr1d_contig_static_kernel(const __grid_constant__ CUtensorMap tmap,
                            const __grid_constant__ Corr1dTmaParams<float, TS> p)
{
    constexpr bool kBox = sizeof(TS) != sizeof(float);   // exact integer box filter
    constexpr int EB = 16 / (int)sizeof(TS);             // elements per 16-byte shared load
    constexpr int R = 16, PC = 4 * R, LA = NT;
    constexpr int NF = (R + NT + EB - 1) / EB * EB;       // row window: whole 16-byte blocks
    extern __shared__ __align__(128) unsigned char smem_raw[];
    TS *tile = aligned_smem<TS>(smem_raw);
    __shared__ __align__(8) uint64_t full_bar;

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int P = p.pitch;
    const int TR = p.tile_rows;
    const int SX = p.ncg * PC;
    const int nrg = TR >> 3;
    const int64_t nrows = p.outer;

    const int64_t t = blockIdx.x;
    const int tcx = (int)(t % p.c_tiles);
    const int64_t trow = t / p.c_tiles;
    const int64_t x0t = (int64_t)tcx * SX;
    const int64_t row0 = trow * TR;
    const int64_t g0 = x0t - p.lo;

    contig_stage_tile<TS, 8>(tmap, tile, &full_bar, p.in, g0, row0, p.n, nrows, P, TR, p.mode, p.cval);

    // a warp keeps one 64-column group and walks the 8-row groups (uniform trip
    // count); CTAs have 8 or 4 warps
    const int cg_shift = p.ncg == 4 ? 2 : (p.ncg == 2 ? 1 : 0);
    const int cg = warp & (p.ncg - 1);
    const int rg0 = warp >> cg_shift;
    const int rg_step = (int)(blockDim.x >> 5) >> cg_shift;
    const int patch_iters = (nrg + rg_step - 1) / rg_step;
    const int r_in = lane & 7;
    const int xl = cg * PC + (lane >> 3) * R;
    const int64_t x = x0t + xl;
    const TS *const tile_thr = tile + r_in * P + xl;
    TS *const out_thr = p.out + ((row0 + r_in) * p.n + x);
    for (int qi = 0; qi < patch_iters; ++qi) {
        const int rg_raw = rg0 + qi * rg_step;
        const bool patch_live = rg_raw < nrg;
        const int rg = patch_live ? rg_raw : 0;
        const TS *srow = tile_thr + rg * 8 * P;
        float f[NF];
        if constexpr (!kBox) {
#pragma unroll
            for (int b = 0; b < NF / 4; ++b) {
                const float4 v = *reinterpret_cast<const float4 *>(srow + 4 * b);
                f[4 * b] = v.x; f[4 * b + 1] = v.y; f[4 * b + 2] = v.z; f[4 * b + 3] = v.w;
            }
        } else {
            // 8 elements per load (one LDS.128 of uint16: the quarter warp's 8 rows
            // hit 8 different bank groups), widened with PRMT + a packed unbias
#pragma unroll
            for (int b = 0; b < NF / 8; ++b) {
                const Vec<float, 8> v = load_block<float, TS, 8>(srow + 8 * b);
#pragma unroll
                for (int i = 0; i < 8; ++i) f[8 * b + i] = v.v[i];
            }
        }
        float2 e[8], o[9];
#pragma unroll
        for (int i = 0; i < 8; ++i) e[i] = make_float2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 9; ++i) o[i] = make_float2(0.f, 0.f);
#pragma unroll
        for (int tt = 0; tt < LA; ++tt) {
            const float wk = p.w[tt];
            const float2 w2 = make_float2(wk, wk);
            // Alignment padding in front of the taps (up to 3; for the box filter,
            // whose taps are 0 or 1, up to 15 and some behind): a warp-uniform
            // skip.  It keeps non-finite neighbours outside the real window from
            // leaking in as 0 * inf, and is cheaper than 8 dead FFMA2.
            if ((kBox || tt < 3) && wk == 0.f) continue;
            if ((tt & 1) == 0) {
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    e[i] = __ffma2_rn(make_float2(f[2 * i + tt], f[2 * i + tt + 1]), w2, e[i]);
            } else {
#pragma unroll
                for (int i = 0; i < 9; ++i)
                    o[i] = __ffma2_rn(make_float2(f[2 * i - 1 + tt], f[2 * i + tt]), w2, o[i]);
            }
        }
        const int64_t row = row0 + rg * 8 + r_in;
        if (patch_live && row < nrows && x < p.n) {
            TS *dst0 = out_thr + (int64_t)(rg * 8) * p.n;
            float res[R];
#pragma unroll
            for (int j = 0; j < R; ++j)
                res[j] = ((j & 1) ? e[j >> 1].y : e[j >> 1].x) + ((j & 1) ? o[(j + 1) >> 1].x : o[j >> 1].y);
            if constexpr (kBox) {
                // rows are whole multiples of 16 bytes and x is a multiple of 16 elements
                store_box_quotients<float, TS, 8>(dst0, reinterpret_cast<const float(&)[8]>(res[0]), p.qinv, p.qc0);
                if (x + 8 < p.n)
                    store_box_quotients<float, TS, 8>(dst0 + 8, reinterpret_cast<const float(&)[8]>(res[8]), p.qinv,
                                                      p.qc0);
            } else if (p.wide_store && x + R <= p.n) {
                if (p.epilogue != B200_EPI_STORE) {
                    const bool reads_out = epilogue_reads_out(p.epilogue);
                    float prev[R];
                    if (reads_out) {
                        ld_global_256(dst0, reinterpret_cast<float(&)[8]>(prev[0]));
                        ld_global_256(dst0 + 8, reinterpret_cast<float(&)[8]>(prev[8]));
                    }
#pragma unroll
                    for (int j = 0; j < R; ++j)
                        res[j] = apply_epilogue<float>(p.epilogue, reads_out ? prev[j] : 0.f, res[j]);
                }
                st_global_256(dst0, reinterpret_cast<const float(&)[8]>(res[0]));
                st_global_256(dst0 + 8, reinterpret_cast<const float(&)[8]>(res[8]));
            } else {
#pragma unroll
                for (int h = 0; h < 4; ++h) {
                    if (x + h * 4 < p.n) {   // rows are whole multiples of 16 bytes
                        float4 *dst = reinterpret_cast<float4 *>(dst0 + h * 4);
                        float4 r4 = make_float4(res[4 * h], res[4 * h + 1], res[4 * h + 2], res[4 * h + 3]);
                        if (p.epilogue != B200_EPI_STORE) {
                            const bool reads_out = epilogue_reads_out(p.epilogue);
                            float4 pv = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (reads_out) pv = *dst;
                            r4.x = apply_epilogue<float>(p.epilogue, pv.x, r4.x);
                            r4.y = apply_epilogue<float>(p.epilogue, pv.y, r4.y);
                            r4.z = apply_epilogue<float>(p.epilogue, pv.z, r4.z);
                            r4.w = apply_epilogue<float>(p.epilogue, pv.w, r4.w);
                        }
                        *dst = r4;
                    }
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------
// What the launchers need beyond Corr1dArgs when the storage type differs from
// the arithmetic type (integer box filter): taps are 0/1 floats, the epilogue
// divides by `size`.
struct BoxInfo {
    bool on = false;
    float qinv = 0.f, qc0 = 0.f;
};

template <typename T, typename TS>
static const char *kernel_name(bool strided)
{
    if (sizeof(TS) != sizeof(T))
        return strided ? (sizeof(TS) == 2 ? "uniform1d_tma_strided_u16" : "uniform1d_tma_strided_u8")
                       : (sizeof(TS) == 2 ? "uniform1d_tma_contig_u16" : "uniform1d_tma_contig_u8");
    return strided ? (sizeof(T) == 4 ? "corr1d_tma_strided_f32" : "corr1d_tma_strided_f64")
                   : (sizeof(T) == 4 ? "corr1d_tma_contig_f32" : "corr1d_tma_contig_f64");
}

template <typename T, typename TS, int R, int MINB, int REM, int TY = 8, int OP = 0>
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
    uint64_t dims[3] = {(uint64_t)g.inner, (uint64_t)n_tot, (uint64_t)g.outer};
    uint64_t strides[2] = {(uint64_t)g.inner * es, (uint64_t)g.inner * es * (uint64_t)n_tot};
    uint32_t bx[3] = {(uint32_t)W, (uint32_t)smem_rows, 1u};
    if (encode_tensor_map(&tmap, dtype_code<TS>(), (void *)base, 3, dims, strides, bx) != 0) return -1;

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
    if constexpr (OP == 0) {
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
    note_kernel(OP == 0 ? kernel_name<T, TS>(true) : (OP == 1 ? "min1d_tma_strided" : "max1d_tma_strided"));
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
    if ((g.inner * es) % 16 != 0) return -1;
    const int64_t n_tot = a.pad_lo + g.n + a.pad_hi;
    if (g.inner >= (1ll << 31) || n_tot >= (1ll << 31) || g.outer >= (1ll << 31)) return -1;
    if ((a.pad_lo || a.pad_hi) && g.outer != 1) return -1;   // pads exist only for axis 0
    const char *base = (const char *)a.in - a.pad_lo * g.inner * es;
    if (((uintptr_t)base & 15) || ((uintptr_t)a.out & 15)) return -1;

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
    if constexpr (sizeof(T) == sizeof(TS)) {
        if (a.g.n <= 8) return launch_strided_r<T, TS, 1>(a, box);
        if (a.g.n <= 16) return launch_strided_r<T, TS, 2>(a, box);
    }
    return launch_strided_r<T, TS, 4>(a, box);
}


template <typename T, typename TS, int SH, int REM>
static int launch_contig_cfg(const Corr1dArgs &a, const BoxInfo &box, int lo_al, int L)
{
    constexpr int V = StorageV<T, TS>::V;
    constexpr int R = 2 * V, PC = 4 * R;
    constexpr int U16B = 16 / (int)sizeof(TS);   // elements per 16 bytes of storage
    const int es = (int)sizeof(TS);
    const LineGeom &g = a.g;
    // tile width: ncg column groups of 4*R outputs; pitch = smallest odd multiple
    // of 16 bytes covering SX + L + 2V - 1 elements (the float path's odd-offset
    // pairs read one block further); TMA boxes are at most 256 elements wide.
    auto pitch_for = [&](int cand) {
        int pt = cand * PC + L + 2 * V - 1;
        pt = (pt + U16B - 1) / U16B;   // 16-byte units
        if (!(pt & 1)) ++pt;
        return pt * U16B;
    };
    // Geometry from the B200 sweeps (profiles/): ~42 KB tiles so that 4 CTAs stay
    // resident per SM; 2 column groups for short filters (taller tiles), 4 for
    // long ones (less halo re-read per output).
    int ncg = env_int("B200_TMA_NCG", L <= 24 ? 2 : 4);
    ncg = ncg >= 4 ? 4 : (ncg >= 2 ? 2 : 1);   // a power of two: warps own whole column groups
    while (ncg > 1 && pitch_for(ncg) > 256) ncg >>= 1;
    const int pitch = pitch_for(ncg);
    if (pitch > 256) return -1;
    const int SX = ncg * PC;
    int tr = env_int("B200_TMA_TR", 0);
    if (tr <= 0) tr = (int)(43 * 1024 / ((size_t)pitch * es));
    tr = tr / 8 * 8;
    tr = tr < 8 ? 8 : (tr > 256 ? 256 : tr);
    while (tr > 8 && (size_t)tr * pitch * es + 512 > kSmemPerSm - 3072) tr -= 8;
    const int64_t rows8 = (g.outer + 7) / 8 * 8;
    if (tr > rows8) tr = (int)rows8;
    const int64_t c_tiles = (g.n + SX - 1) / SX;
    const int64_t r_tiles = (g.outer + tr - 1) / tr;
    const int64_t num_tiles = c_tiles * r_tiles;
    if (num_tiles >= (1ll << 31)) return -1;
    const size_t stage_bytes = (size_t)tr * pitch * es;

    CUtensorMap tmap;
    uint64_t dims[2] = {(uint64_t)g.n, (uint64_t)g.outer};
    uint64_t strides[1] = {(uint64_t)g.n * es};
    uint32_t bx[2] = {(uint32_t)pitch, (uint32_t)tr};
    if (encode_tensor_map(&tmap, dtype_code<TS>(), (void *)a.in, 2, dims, strides, bx) != 0) return -1;

    Corr1dTmaParams<T, TS> p;
    p.in = (const TS *)a.in; p.out = (TS *)a.out;
    p.outer = g.outer; p.n = g.n; p.inner = 1;
    p.pad_lo = 0; p.pad_hi = 0;
    p.nw = L; p.lo = lo_al; p.mode = a.mode; p.epilogue = a.epilogue;
    p.tile_rows = tr; p.smem_rows = tr; p.nb = 1; p.pitch = pitch; p.shift = SH; p.ncg = ncg;
    p.wide_store = (((g.n * es) % 32) == 0 && ((uintptr_t)a.out & 31) == 0) ? 1 : 0;
    p.n_tiles = (int)r_tiles; p.c_tiles = (int)c_tiles;
    p.cval = (TS)a.cval;
    p.qinv = (T)box.qinv; p.qc0 = (T)box.qc0;
    // taps: `L - a.nw` leading zeros (alignment of the TMA box start), then the
    // caller's; the integer box filter pads to a multiple of V with zeros after
    const int lead = lo_al - (a.nw / 2 + a.origin);
    for (int i = 0; i < L; ++i) p.w[i] = T(0);
    for (int i = 0; i < a.nw; ++i) p.w[lead + i] = (T)a.w[i];

    auto kern = corr1d_contig_tma_kernel<T, TS, SH, REM>;
    const size_t smem = stage_bytes + 128;
    B200_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemOptIn));
    kern<<<(unsigned)num_tiles, 256, smem, a.stream>>>(tmap, p);
    B200_CUDA_TRY(cudaGetLastError());
    count_launch();
    note_kernel(kernel_name<T, TS>(false));
    return B200_OK;
}

template <typename TS, int NT>
static int launch_contig_static_cfg(const Corr1dArgs &a, const BoxInfo &box, int lo_al, int lead)
{
    constexpr int EB = 16 / (int)sizeof(TS);   // elements per 16 bytes of storage
    constexpr int R = 16, PC = 4 * R;
    constexpr int LA = (R + NT + EB - 1) / EB * EB - R;   // window beyond the 16 outputs
    const int es = (int)sizeof(TS);
    const LineGeom &g = a.g;
    auto pitch_for = [&](int cand) {
        int pu = (cand * PC + LA + EB - 1) / EB;
        if (!(pu & 1)) ++pu;
        return pu * EB;
    };
    // Geometry from the B200 sweep (tools/sweep_static2.py, profiles/): 4-warp
    // CTAs, many of them resident, so that one CTA's load phase hides behind the
    // others' FMA phases; two 64-column groups for short filters (less halo per
    // output), one for long ones (smaller tiles).
    const int threads = env_int("B200_TMA_STHREADS", 128);
    const int ctas = env_int("B200_TMA_SCTAS", NT < 24 ? 8 : (NT < 32 ? 6 : 4));
    int ncg = env_int("B200_TMA_SNCG", (NT < 24 || (NT >= 32 && NT < 40)) ? 2 : 1);
    ncg = ncg >= 2 ? 2 : 1;
    if (pitch_for(ncg) > 256) ncg = 1;
    const int pitch = pitch_for(ncg);
    if (pitch > 256) return -1;
    const int SX = ncg * PC;
    int tr = env_int("B200_TMA_STR", 0);
    if (tr <= 0) tr = (int)(((kSmemPerSm - 3072) / ctas - 256) / ((size_t)pitch * es));
    tr = tr / 8 * 8;
    tr = tr < 8 ? 8 : (tr > 256 ? 256 : tr);
    const int64_t rows8 = (g.outer + 7) / 8 * 8;
    if (tr > rows8) tr = (int)rows8;
    const int64_t c_tiles = (g.n + SX - 1) / SX;
    const int64_t r_tiles = (g.outer + tr - 1) / tr;
    const int64_t num_tiles = c_tiles * r_tiles;
    if (num_tiles >= (1ll << 31)) return -1;
    const size_t stage_bytes = (size_t)tr * pitch * es;

    CUtensorMap tmap;
    uint64_t dims[2] = {(uint64_t)g.n, (uint64_t)g.outer};
    uint64_t strides[1] = {(uint64_t)g.n * es};
    uint32_t bx[2] = {(uint32_t)pitch, (uint32_t)tr};
    if (encode_tensor_map(&tmap, dtype_code<TS>(), (void *)a.in, 2, dims, strides, bx) != 0) return -1;

    Corr1dTmaParams<float, TS> p;
    p.in = (const TS *)a.in; p.out = (TS *)a.out;
    p.outer = g.outer; p.n = g.n; p.inner = 1;
    p.pad_lo = 0; p.pad_hi = 0;
    p.nw = NT; p.lo = lo_al; p.mode = a.mode; p.epilogue = a.epilogue;
    p.tile_rows = tr; p.smem_rows = tr; p.nb = 1; p.pitch = pitch; p.shift = 0; p.ncg = ncg;
    p.wide_store = (((g.n * es) % 32) == 0 && ((uintptr_t)a.out & 31) == 0) ? 1 : 0;
    p.n_tiles = (int)r_tiles; p.c_tiles = (int)c_tiles;
    p.cval = (TS)a.cval;
    p.qinv = box.qinv; p.qc0 = box.qc0;
    for (int i = 0; i < NT; ++i) p.w[i] = 0.f;
    for (int i = 0; i < a.nw; ++i) p.w[lead + i] = (float)a.w[i];

    auto kern = corr1d_contig_static_kernel<TS, NT>;
    const size_t smem = stage_bytes + 128;
    B200_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemOptIn));
    kern<<<(unsigned)num_tiles, threads == 128 ? 128 : 256, smem, a.stream>>>(tmap, p);
    B200_CUDA_TRY(cudaGetLastError());
    count_launch();
    note_kernel(kernel_name<float, TS>(false));
    return B200_OK;
}

// float32 rows with at most kStaticTaps taps (alignment zeros in front
// included): the compile-time tap engine.  Returns -2 when the filter is longer.
constexpr int kStaticTaps = 48;
static int launch_contig_static(const Corr1dArgs &a)
{
    const LineGeom &g = a.g;
    if (a.pad_lo || a.pad_hi) return -1;
    if ((g.n * 4) % 16 != 0) return -1;
    if (g.n >= (1ll << 31) || g.outer >= (1ll << 31)) return -1;
    if (((uintptr_t)a.in & 15) || ((uintptr_t)a.out & 15)) return -1;
    const int lo = a.nw / 2 + a.origin;
    const int lo_al = (lo + 3) / 4 * 4;
    const int lead = lo_al - lo;
    const int nt = lead + a.nw;
    if (nt > kStaticTaps) return -2;
    const BoxInfo none;
    switch (nt) {
#define B200_C1_CASE(K) case K: return launch_contig_static_cfg<float, K>(a, none, lo_al, lead);
#define B200_C1_CASE4(K) B200_C1_CASE(K) B200_C1_CASE(K + 1) B200_C1_CASE(K + 2) B200_C1_CASE(K + 3)
    B200_C1_CASE4(1) B200_C1_CASE4(5) B200_C1_CASE4(9) B200_C1_CASE4(13) B200_C1_CASE4(17) B200_C1_CASE4(21)
    B200_C1_CASE4(25) B200_C1_CASE4(29) B200_C1_CASE4(33) B200_C1_CASE4(37) B200_C1_CASE4(41) B200_C1_CASE4(45)
#undef B200_C1_CASE4
#undef B200_C1_CASE
    default: return -2;
    }
}

template <typename T, int SH>
static int launch_contig_sh(const Corr1dArgs &a, int lo_al, int L)
{
    constexpr int V = 16 / (int)sizeof(T);
    const BoxInfo none;
    switch (L % V) {
    case 0: return launch_contig_cfg<T, T, SH, 0>(a, none, lo_al, L);
    case 1: return launch_contig_cfg<T, T, SH, 1>(a, none, lo_al, L);
    case 2: if constexpr (V == 4) return launch_contig_cfg<T, T, SH, 2>(a, none, lo_al, L);
    case 3: if constexpr (V == 4) return launch_contig_cfg<T, T, SH, 3>(a, none, lo_al, L);
    }
    return -1;
}

template <typename T>
static int launch_contig(const Corr1dArgs &a)
{
    constexpr int V = 16 / (int)sizeof(T);
    const int es = (int)sizeof(T);
    const LineGeom &g = a.g;
    if (a.pad_lo || a.pad_hi) return -1;
    if ((g.n * es) % 16 != 0) return -1;
    if (g.n >= (1ll << 31) || g.outer >= (1ll << 31)) return -1;
    if (((uintptr_t)a.in & 15) || ((uintptr_t)a.out & 15)) return -1;
    // TMA needs the box's innermost start coordinate on a 16-byte boundary
    // (cp.async.bulk.tensor faults otherwise): start the box `shift` elements
    // early and skip `shift` leading taps in the kernel instead.
    const int lo = a.nw / 2 + a.origin;
    const int lo_al = (lo + V - 1) / V * V;
    const int shift = lo_al - lo;
    const int L = a.nw + shift;
    if (L > kMaxTapsTma) return -1;
    switch (shift) {
    case 0: return launch_contig_sh<T, 0>(a, lo_al, L);
    case 1: return launch_contig_sh<T, 1>(a, lo_al, L);
    case 2: if constexpr (V == 4) return launch_contig_sh<T, 2>(a, lo_al, L);
    case 3: if constexpr (V == 4) return launch_contig_sh<T, 3>(a, lo_al, L);
    }
    return -1;
}

// integer box filter along the contiguous axis: the TMA box start must sit on
// a 16-byte boundary of the STORAGE type (8 u16 / 16 u8 elements), so up to 15
// zero taps lead; zero taps also pad the tail to a whole group (SH = REM = 0:
// one kernel variant per storage type)
template <typename TS>
static int launch_contig_box(const Corr1dArgs &a, const BoxInfo &box)
{
    constexpr int V = 4;
    constexpr int U16B = 16 / (int)sizeof(TS);
    const int es = (int)sizeof(TS);
    const LineGeom &g = a.g;
    if (a.pad_lo || a.pad_hi) return -1;
    if ((g.n * es) % 16 != 0) return -1;
    if (g.n >= (1ll << 31) || g.outer >= (1ll << 31)) return -1;
    if (((uintptr_t)a.in & 15) || ((uintptr_t)a.out & 15)) return -1;
    const int lo = a.nw / 2 + a.origin;
    const int lo_al = (lo + U16B - 1) / U16B * U16B;
    const int lead = lo_al - lo;
    if (env_int("B200_TMA_STATIC", 1)) {
        // compile-time tap engine: taps padded to whole 16-byte blocks of storage
        const int la = (lead + a.nw + U16B - 1) / U16B * U16B;
        switch (la) {
        case 8: if constexpr (U16B == 8) return launch_contig_static_cfg<TS, 8>(a, box, lo_al, lead); break;
        case 16: return launch_contig_static_cfg<TS, 16>(a, box, lo_al, lead);
        case 24: if constexpr (U16B == 8) return launch_contig_static_cfg<TS, 24>(a, box, lo_al, lead); break;
        case 32: return launch_contig_static_cfg<TS, 32>(a, box, lo_al, lead);
        case 40: if constexpr (U16B == 8) return launch_contig_static_cfg<TS, 40>(a, box, lo_al, lead); break;
        case 48: return launch_contig_static_cfg<TS, 48>(a, box, lo_al, lead);
        default: break;
        }
    }
    int L = a.nw + lead;
    L = (L + V - 1) / V * V;
    if (L > kMaxTapsTma) return -1;
    return launch_contig_cfg<float, TS, 0, 0>(a, box, lo_al, L);
}

int launch_corr1d_tma(const Corr1dArgs &a)
{
    // NB the choice of kernel family must not depend on the number of planes
    // of the call: slabs of one volume (any GPU count) have to take the same
    // arithmetic path so that the whole-volume result is partition-invariant.
    const BoxInfo none;
    if (a.dtype == B200_F32) {
        if (a.g.inner != 1) return launch_strided<float, float>(a, none);
        // (which contiguous-axis engine runs depends on the tap count only)
        const int rc = env_int("B200_TMA_STATIC", 1) ? launch_contig_static(a) : -2;
        return rc == -2 ? launch_contig<float>(a) : rc;
    }
    if (a.dtype == B200_F64)
        return a.g.inner == 1 ? launch_contig<double>(a) : launch_strided<double, double>(a, none);
    return -1;
}

// ---------------------------------------------------------------------------
// uniform_filter1d fast paths
// ---------------------------------------------------------------------------
// Host proof that the float epilogue reproduces the reference's
// trunc((double)sum / size) for EVERY window sum 0 <= s <= vmax * size: the
// rounded value is monotone in s, so it is enough that the first and the last
// sum of every quotient class [q*size, q*size + size - 1] land on q.
static bool uniform_int_exact(int size, int64_t vmax)
{
    if (size < 2 || vmax * size >= (1ll << 23)) return false;
    const float qinv = 1.0f / (float)size;
    const float qc0 = 0.5f - 0.5f * (float)size;
    auto quot = [&](int64_t s) -> int64_t {
        const float y = fmaf((float)s + qc0, qinv, 12582912.f);
        uint32_t bits;
        memcpy(&bits, &y, 4);
        return (int64_t)(bits & 0x3FFFFFu);
    };
    for (int64_t q = 0; q <= vmax; ++q) {
        const int64_t s0 = q * size, s1 = q * size + size - 1;
        const int64_t want = (int64_t)((double)s0 / (double)size);
        if (quot(s0) != want) return false;
        if (s1 <= vmax * size && quot(s1) != (int64_t)((double)s1 / (double)size)) return false;
    }
    return true;
}

// can the exact float epilogue serve this integer box filter?
static bool uniform_box_ok(int dtype, int size, int mode, double cval)
{
    if (dtype != B200_U16 && dtype != B200_U8) return false;
    const int64_t vmax = dtype == B200_U16 ? 65535 : 255;
    if (size < 2 || size > 128) return false;
    if (mode == B200_MODE_CONSTANT) {
        // the tile holds storage-type values: cval must be one
        if (!(cval >= 0.0 && cval <= (double)vmax) || cval != (double)(int64_t)cval) return false;
    }
    // (one proof per size and type; sizes are small, the loop is ~vmax iterations)
    static std::mutex mu;
    static signed char proven[2][129];   // 0 unknown, 1 exact, -1 not
    const int ti = dtype == B200_U16 ? 1 : 0;
    std::lock_guard<std::mutex> lk(mu);
    if (proven[ti][size] == 0) proven[ti][size] = uniform_int_exact(size, vmax) ? 1 : -1;
    return proven[ti][size] > 0;
}

template <typename TS>
static int launch_uniform_box(const Uniform1dArgs &u, int64_t vmax)
{
    (void)vmax;
    if (!uniform_box_ok(u.dtype, u.size, u.mode, u.cval)) return -1;
    double ones[128];
    for (int i = 0; i < u.size; ++i) ones[i] = 1.0;
    Corr1dArgs a;
    a.in = u.in; a.out = u.out; a.dtype = u.dtype; a.g = u.g;
    a.w = ones; a.nw = u.size; a.mode = u.mode; a.cval = u.cval; a.origin = u.origin;
    a.pad_lo = u.pad_lo; a.pad_hi = u.pad_hi; a.epilogue = B200_EPI_STORE; a.flags = u.flags;
    a.stream = u.stream;
    BoxInfo box;
    box.on = true;
    box.qinv = 1.0f / (float)u.size;
    box.qc0 = 0.5f - 0.5f * (float)u.size;
    return u.g.inner == 1 ? launch_contig_box<TS>(a, box) : launch_strided<float, TS>(a, box);
}

int launch_uniform1d_fast(const Uniform1dArgs &u)
{
    if (u.size < 2) return -1;
    if (u.dtype == B200_U16) return launch_uniform_box<uint16_t>(u, 65535);
    if (u.dtype == B200_U8) return launch_uniform_box<uint8_t>(u, 255);
    if (u.dtype == B200_F32 || u.dtype == B200_F64) {
        // floats: the box filter is a correlation with `size` taps of 1/size
        // (fp32 / fp64 accumulation in the array's type, like every float pass)
        if (u.size > kMaxTapsTma) return -1;
        double w[kMaxTapsTma];
        for (int i = 0; i < u.size; ++i) w[i] = 1.0 / (double)u.size;
        Corr1dArgs a;
        a.in = u.in; a.out = u.out; a.dtype = u.dtype; a.g = u.g;
        a.w = w; a.nw = u.size; a.mode = u.mode; a.cval = u.cval; a.origin = u.origin;
        a.pad_lo = u.pad_lo; a.pad_hi = u.pad_hi; a.epilogue = B200_EPI_STORE; a.flags = u.flags;
        a.stream = u.stream;
        return launch_corr1d_tma(a);
    }
    return -1;
}

// box filter along axis 0 of a slab, neighbour planes read in place
int uniform1d_peer_eligible(int dtype, int64_t n, int64_t inner, int size, int origin, int mode, double cval)
{
    if (size < 2 || size / 2 + origin < 0 || size / 2 + origin >= size) return 0;
    if (dtype == B200_F32 || dtype == B200_F64) return corr1d_peer_eligible(dtype, n, inner, size, origin);
    if (!uniform_box_ok(dtype, size, mode, cval)) return 0;
    return corr1d_peer_eligible(dtype, n, inner, size, origin);
}

int launch_uniform1d_peer(const Uniform1dArgs &u, const void *peer_lo, int n_lo, const void *peer_hi, int n_hi)
{
    if (u.size < 2 || u.size > kMaxTapsTma) return -1;
    double w[kMaxTapsTma];
    const bool is_float = u.dtype == B200_F32 || u.dtype == B200_F64;
    if (!is_float && !uniform_box_ok(u.dtype, u.size, u.mode, u.cval)) return -1;
    for (int i = 0; i < u.size; ++i) w[i] = is_float ? 1.0 / (double)u.size : 1.0;
    Corr1dArgs a;
    a.in = u.in; a.out = u.out; a.dtype = u.dtype; a.g = u.g;
    a.w = w; a.nw = u.size; a.mode = u.mode; a.cval = u.cval; a.origin = u.origin;
    a.pad_lo = 0; a.pad_hi = 0; a.epilogue = B200_EPI_STORE; a.flags = u.flags; a.stream = u.stream;
    a.peer_lo = peer_lo; a.peer_hi = peer_hi; a.n_peer_lo = n_lo; a.n_peer_hi = n_hi;
    if (is_float) return launch_corr1d_peer(a);
    BoxInfo box;
    box.on = true;
    box.qinv = 1.0f / (float)u.size;
    box.qc0 = 0.5f - 0.5f * (float)u.size;
    return u.dtype == B200_U16 ? launch_peer_typed<float, uint16_t>(a, box) : launch_peer_typed<float, uint8_t>(a, box);
}


// ===========================================================================
// minimum / maximum filter along the contiguous axis (float32, uint8/uint16
// storage): the tile, the lane mapping and the stores of the static
// contiguous-axis engine; NT = window taps incl. alignment padding (flagged 0 in
// p.w and skipped), every real tap is 16 FMNMX.
// ===========================================================================
template <typename TS, int NT, bool IS_MAX>
__global__ void __launch_bounds__(256, 3)
minmax_contig_static_kernel(const __grid_constant__ CUtensorMap tmap,
                            const __grid_constant__ Corr1dTmaParams<float, TS> p)
{
    constexpr bool kInt = sizeof(TS) != sizeof(float);
    constexpr int EB = 16 / (int)sizeof(TS);
    constexpr int R = 16, PC = 4 * R;
    constexpr int NF = (R + NT + EB - 1) / EB * EB;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    TS *tile = aligned_smem<TS>(smem_raw);
    __shared__ __align__(8) uint64_t full_bar;

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int P = p.pitch;
    const int TR = p.tile_rows;
    const int SX = p.ncg * PC;
    const int nrg = TR >> 3;
    const int64_t nrows = p.outer;
    const int64_t t = blockIdx.x;
    const int tcx = (int)(t % p.c_tiles);
    const int64_t trow = t / p.c_tiles;
    const int64_t x0t = (int64_t)tcx * SX;
    const int64_t row0 = trow * TR;
    const int64_t g0 = x0t - p.lo;

    contig_stage_tile<TS, 8>(tmap, tile, &full_bar, p.in, g0, row0, p.n, nrows, P, TR, p.mode, p.cval);

    const int cg_shift = p.ncg == 4 ? 2 : (p.ncg == 2 ? 1 : 0);
    const int cg = warp & (p.ncg - 1);
    const int rg0 = warp >> cg_shift;
    const int rg_step = (int)(blockDim.x >> 5) >> cg_shift;
    const int patch_iters = (nrg + rg_step - 1) / rg_step;
    const int r_in = lane & 7;
    const int xl = cg * PC + (lane >> 3) * R;
    const int64_t x = x0t + xl;
    const TS *const tile_thr = tile + r_in * P + xl;
    TS *const out_thr = p.out + ((row0 + r_in) * p.n + x);
    for (int qi = 0; qi < patch_iters; ++qi) {
        const int rg_raw = rg0 + qi * rg_step;
        const bool patch_live = rg_raw < nrg;
        const int rg = patch_live ? rg_raw : 0;
        const TS *srow = tile_thr + rg * 8 * P;
        float f[NF];
        if constexpr (!kInt) {
#pragma unroll
            for (int b = 0; b < NF / 4; ++b) {
                const float4 v = *reinterpret_cast<const float4 *>(srow + 4 * b);
                f[4 * b] = v.x; f[4 * b + 1] = v.y; f[4 * b + 2] = v.z; f[4 * b + 3] = v.w;
            }
        } else {
#pragma unroll
            for (int b = 0; b < NF / 8; ++b) {
                const Vec<float, 8> v = load_block<float, TS, 8>(srow + 8 * b);
#pragma unroll
                for (int i = 0; i < 8; ++i) f[8 * b + i] = v.v[i];
            }
        }
        float m[R];
#pragma unroll
        for (int j = 0; j < R; ++j) m[j] = IS_MAX ? -INFINITY : INFINITY;
#pragma unroll
        for (int tt = 0; tt < NT; ++tt) {
            if (p.w[tt] == 0.f) continue;     // alignment padding: warp-uniform skip
#pragma unroll
            for (int j = 0; j < R; ++j) m[j] = IS_MAX ? fmaxf(m[j], f[j + tt]) : fminf(m[j], f[j + tt]);
        }
        const int64_t row = row0 + rg * 8 + r_in;
        if (patch_live && row < nrows && x < p.n) {
            TS *dst0 = out_thr + (int64_t)(rg * 8) * p.n;
            if constexpr (kInt) {
                // exact small integers in floats: v + 1.5 * 2^23 has v in its low mantissa bits
                store_box_quotients<float, TS, 8>(dst0, reinterpret_cast<const float(&)[8]>(m[0]), 1.f, 0.f);
                if (x + 8 < p.n)
                    store_box_quotients<float, TS, 8>(dst0 + 8, reinterpret_cast<const float(&)[8]>(m[8]), 1.f, 0.f);
            } else if (p.wide_store && x + R <= p.n) {
                st_global_256(dst0, reinterpret_cast<const float(&)[8]>(m[0]));
                st_global_256(dst0 + 8, reinterpret_cast<const float(&)[8]>(m[8]));
            } else {
#pragma unroll
                for (int h = 0; h < 4; ++h)
                    if (x + h * 4 < p.n)
                        *reinterpret_cast<float4 *>(dst0 + h * 4) =
                            make_float4(m[4 * h], m[4 * h + 1], m[4 * h + 2], m[4 * h + 3]);
            }
        }
    }
}

template <typename TS, int NT, bool IS_MAX>
static int launch_minmax_contig_cfg(const Corr1dArgs &a, int lo_al, int lead)
{
    constexpr int EB = 16 / (int)sizeof(TS);
    constexpr int R = 16, PC = 4 * R;
    constexpr int LA = (R + NT + EB - 1) / EB * EB - R;
    const int es = (int)sizeof(TS);
    const LineGeom &g = a.g;
    auto pitch_for = [&](int cand) {
        int pu = (cand * PC + LA + EB - 1) / EB;
        if (!(pu & 1)) ++pu;
        return pu * EB;
    };
    int ncg = 2;
    if (pitch_for(ncg) > 256) ncg = 1;
    const int pitch = pitch_for(ncg);
    if (pitch > 256) return -1;
    const int SX = ncg * PC;
    int tr = (int)(((kSmemPerSm - 3072) / 6 - 256) / ((size_t)pitch * es));
    tr = tr / 8 * 8;
    tr = tr < 8 ? 8 : (tr > 256 ? 256 : tr);
    const int64_t rows8 = (g.outer + 7) / 8 * 8;
    if (tr > rows8) tr = (int)rows8;
    const int64_t c_tiles = (g.n + SX - 1) / SX;
    const int64_t r_tiles = (g.outer + tr - 1) / tr;
    const int64_t num_tiles = c_tiles * r_tiles;
    if (num_tiles >= (1ll << 31)) return -1;
    const size_t stage_bytes = (size_t)tr * pitch * es;

    CUtensorMap tmap;
    uint64_t dims[2] = {(uint64_t)g.n, (uint64_t)g.outer};
    uint64_t strides[1] = {(uint64_t)g.n * es};
    uint32_t bx[2] = {(uint32_t)pitch, (uint32_t)tr};
    if (encode_tensor_map(&tmap, dtype_code<TS>(), (void *)a.in, 2, dims, strides, bx) != 0) return -1;

    Corr1dTmaParams<float, TS> p;
    p.in = (const TS *)a.in; p.out = (TS *)a.out;
    p.outer = g.outer; p.n = g.n; p.inner = 1;
    p.pad_lo = 0; p.pad_hi = 0;
    p.nw = NT; p.lo = lo_al; p.mode = a.mode; p.epilogue = B200_EPI_STORE;
    p.tile_rows = tr; p.smem_rows = tr; p.nb = 1; p.pitch = pitch; p.shift = 0; p.ncg = ncg;
    p.wide_store = (((g.n * es) % 32) == 0 && ((uintptr_t)a.out & 31) == 0) ? 1 : 0;
    p.n_tiles = (int)r_tiles; p.c_tiles = (int)c_tiles;
    p.cval = (TS)a.cval;
    p.qinv = 1.f; p.qc0 = 0.f;
    for (int i = 0; i < NT; ++i) p.w[i] = (i >= lead && i < lead + a.nw) ? 1.f : 0.f;

    auto kern = minmax_contig_static_kernel<TS, NT, IS_MAX>;
    const size_t smem = stage_bytes + 128;
    B200_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemOptIn));
    kern<<<(unsigned)num_tiles, 128, smem, a.stream>>>(tmap, p);
    B200_CUDA_TRY(cudaGetLastError());
    count_launch();
    note_kernel(IS_MAX ? "max1d_tma_contig" : "min1d_tma_contig");
    return B200_OK;
}

template <typename TS, bool IS_MAX>
static int launch_minmax_contig(const Corr1dArgs &a)
{
    constexpr int EB = 16 / (int)sizeof(TS);
    const int es = (int)sizeof(TS);
    const LineGeom &g = a.g;
    if (a.pad_lo || a.pad_hi) return -1;
    if ((g.n * es) % 16 != 0) return -1;
    if (g.n >= (1ll << 31) || g.outer >= (1ll << 31)) return -1;
    if (((uintptr_t)a.in & 15) || ((uintptr_t)a.out & 15)) return -1;
    const int lo = a.nw / 2 + a.origin;
    const int lo_al = (lo + EB - 1) / EB * EB;
    const int lead = lo_al - lo;
    const int nt = lead + a.nw;
    // window lengths are rounded up to the instantiated tap counts (padding taps are skipped)
    if (nt <= 8) return launch_minmax_contig_cfg<TS, 8, IS_MAX>(a, lo_al, lead);
    if (nt <= 16) return launch_minmax_contig_cfg<TS, 16, IS_MAX>(a, lo_al, lead);
    if (nt <= 32) return launch_minmax_contig_cfg<TS, 32, IS_MAX>(a, lo_al, lead);
    if (nt <= 48) return launch_minmax_contig_cfg<TS, 48, IS_MAX>(a, lo_al, lead);
    return -1;
}

template <typename TS>
static int launch_minmax_typed(const MinMax1dArgs &u)
{
    double ones[kMaxTapsTma];
    if (u.size > 64) return -1;
    for (int i = 0; i < u.size; ++i) ones[i] = 1.0;
    Corr1dArgs a;
    a.in = u.in; a.out = u.out; a.dtype = u.dtype; a.g = u.g;
    a.w = ones; a.nw = u.size; a.mode = u.mode; a.cval = u.cval; a.origin = u.origin;
    a.pad_lo = u.pad_lo; a.pad_hi = u.pad_hi; a.epilogue = B200_EPI_STORE; a.flags = u.flags;
    a.stream = u.stream;
    BoxInfo box;
    box.on = true; box.qinv = 1.f; box.qc0 = 0.f;
    if (u.g.inner == 1)
        return u.is_max ? launch_minmax_contig<TS, true>(a) : launch_minmax_contig<TS, false>(a);
    return u.is_max ? launch_strided_r<float, TS, 4, 2>(a, box) : launch_strided_r<float, TS, 4, 1>(a, box);
}

// minimum / maximum filter on the tiled kernels: float32 and uint8 / uint16
// (exact in fp32).  The tile holds storage-type values, so a constant-mode cval
// must be one; -1 = not eligible (the generic exact kernel takes the call).
int launch_minmax1d_fast(const MinMax1dArgs &u)
{
    if (u.size < 2) return -1;
    if (u.dtype == B200_F32) {
        if (u.mode == B200_MODE_CONSTANT && (double)(float)u.cval != u.cval) return -1;
        return launch_minmax_typed<float>(u);
    }
    if (u.dtype == B200_U16 || u.dtype == B200_U8) {
        const double vmax = u.dtype == B200_U16 ? 65535.0 : 255.0;
        if (u.mode == B200_MODE_CONSTANT &&
            (!(u.cval >= 0.0 && u.cval <= vmax) || u.cval != (double)(int64_t)u.cval))
            return -1;
        return u.dtype == B200_U16 ? launch_minmax_typed<uint16_t>(u) : launch_minmax_typed<uint8_t>(u);
    }
    return -1;
}

}  // namespace b200
