// corr2d_fused.cu -- TWO consecutive float32 1-D passes, along the last two axes
// of the array, in ONE launch (sm_100a).
//
// scipy's separable filters (scipy/ndimage/_filters.py:852-858) run one
// correlate1d per axis and store every pass; on the GPU each stored pass is a
// full HBM round trip (8 B/voxel for float32).  The pass along axis ndim-2 and
// the pass along axis ndim-1 are fused here: a CTA stages a
// (TYO + L1 - 1) x 256 input tile of one plane (one cp.async.bulk.tensor.3d, or
// plain loads when the rows are not 16-byte multiples), runs the axis ndim-2
// pass out of shared memory into a second shared tile -- in float32, i.e.
// rounded exactly as the HBM store of the unfused pass would round it -- and then
// the contiguous pass out of that tile into HBM.  16 B/voxel for a 3-D gaussian
// instead of 24; a 2-D gaussian is a single launch.
//
// Both phases perform the same operations in the same order as the unfused
// kernels they replace (corr1d_strided.cu: acc = fma(x, w[k], acc), k ascending;
// corr1d_contig.cu: the static engine's even / odd packed-pair sums), so the
// result is bit-identical to the two-launch path.
//
// Phase A (axis ndim-2): a lane owns a 16-byte column vector and RA consecutive
// output rows; every staged row is loaded once (LDS.128) and feeds all RA
// accumulators with RA different weights (weights sit in uniform registers, so
// the window lives in the weights, not in vector registers): 2*RA packed FFMA2
// per LDS.128.  Phase B (axis ndim-1): the compile-time tap engine of
// corr1d_contig_static_kernel (16 outputs per thread, LDS.128 row window, FFMA2
// on aligned pairs), stores are 256-bit.
#include "corr1d_tma.cuh"

namespace b200 {

constexpr int kFusedPX = 256;       // staged tile width in floats (TMA box limit)
constexpr int kFusedPI = 260;       // pitch of the intermediate tile: 16 bytes * odd
constexpr int kFusedMaxTaps = 48;

// floats of one staged buffer when the intermediate tile and the staged results alias the input box
constexpr int fused_buffer_floats(int l1, int tyo, int nwarps)
{
    const int in = (tyo + l1 - 1) * kFusedPX, mid = tyo * kFusedPI;
    (void)nwarps;
    return ((in > mid ? in : mid) + 31) / 32 * 32;
}

struct Corr2dParams {
    const void *in;       // float, or the storage type of the exact integer box filter
    void *out;
    int outer, ny, nx;
    int lo1, mode1;       // rows read above an output row: nw1/2 + origin1
    int lo2, mode2;       // staged column 0 is array column x0 - lo2 (lo2 = reach in front, rounded up to 4)
    int hi1, hi2;         // rows / columns read beyond an output
    int epilogue, wide_store;
    int y_tiles, x_tiles, n_tiles;
    int ncg;              // 16-output column groups per tile in use: tiles are ncg * 16 columns wide
    float cval;
    float w1[kFusedMaxTaps];   // taps along axis ndim-2
    float w2[kFusedMaxTaps];   // taps along axis ndim-1, alignment zeros in front
    float qinv, qc0;      // integer box filter: the quotient epilogue (tile_common.cuh: box_quotient_bits)
#ifdef B200_BOX_DEBUG
    int dbg;              // bring-up: leave the tile after stage `dbg`
#endif
};

#ifdef B200_BOX_DEBUG
#define B200_BOX_STAGE(k) do { if constexpr (BOX) { if (p.dbg == (k)) return; } } while (0)
#else
#define B200_BOX_STAGE(k) do { } while (0)
#endif

// 32-bit form of ext_index() (common.cuh) for tile coordinates: |i| and n are below 2^31 here
__device__ __forceinline__ int ext_index32(int i, int n, int mode)
{
    if (i >= 0 && i < n) return i;
    switch (mode) {
    case B200_MODE_NEAREST:
        return i < 0 ? 0 : n - 1;
    case B200_MODE_WRAP: {
        int t = i % n;
        return t < 0 ? t + n : t;
    }
    case B200_MODE_REFLECT: {
        const long long pp = 2ll * n;
        long long t = i % pp;
        if (t < 0) t += pp;
        return (int)(t < n ? t : pp - 1 - t);
    }
    case B200_MODE_MIRROR: {
        if (n == 1) return 0;
        const long long pp = 2ll * n - 2;
        long long t = i % pp;
        if (t < 0) t += pp;
        return (int)(t < n ? t : pp - t);
    }
    default:
        return -1;
    }
}

// Exact integer box filter (TS = unsigned char / unsigned short): the staged box stays in the storage
// type (half / a quarter of the shared-memory traffic of a float tile -- the float form of this kernel
// ran at 90 % of the shared-memory pipe) and is widened in registers where phase A reads it: an
// integer v < 2^23 is the low mantissa of 2^23 + v, i.e. one byte permute and one exact subtraction
// per element, no I2F.
// MAGIC: leave the elements as 2^23 + v (no subtraction): differences of two such values are exact, which is
// all the running sums need once the first window has been summed
template <typename TS, bool MAGIC = false>
__device__ __forceinline__ float4 box_load4(const TS *ptr)
{
    auto widen = [](uint32_t bits) {
        const float f = __uint_as_float(bits);        // 2^23 + v
        if constexpr (MAGIC) return f;
        else return f - 8388608.f;
    };
    float4 v;
    if constexpr (sizeof(TS) == 2) {
        const uint2 w = *reinterpret_cast<const uint2 *>(ptr);
        v.x = widen(__byte_perm(w.x, 0x4B000000u, 0x7410));
        v.y = widen(__byte_perm(w.x, 0x4B000000u, 0x7432));
        v.z = widen(__byte_perm(w.y, 0x4B000000u, 0x7410));
        v.w = widen(__byte_perm(w.y, 0x4B000000u, 0x7432));
    } else {
        static_assert(sizeof(TS) == 1, "8- or 16-bit storage");
        const uint32_t w = *reinterpret_cast<const uint32_t *>(ptr);
        v.x = widen(__byte_perm(w, 0x4B000000u, 0x7440));
        v.y = widen(__byte_perm(w, 0x4B000000u, 0x7441));
        v.z = widen(__byte_perm(w, 0x4B000000u, 0x7442));
        v.w = widen(__byte_perm(w, 0x4B000000u, 0x7443));
    }
    return v;
}

// bytes of the staged box in the storage type, rounded up so that the float tile behind it stays aligned
template <typename TS>
constexpr int box_raw_bytes(int tyin)
{
    return (tyin * kFusedPX * (int)sizeof(TS) + 127) / 128 * 128;
}

// Edge tiles (and every tile of the plain-load variant): fill the staged elements that lie outside
// the array -- TMA zero-fills them -- through the extension maps of the two axes.  Kept out of line:
// interior CTAs never come here, and the hot path keeps its registers.
template <int TYIN, int NTHR, bool TMA, typename TS>
__device__ __noinline__ void fused_stage_edges(const Corr2dParams &p, TS *tin, int *row_src, int *col_src,
                                               const TS *plane, int gx0, int gy0, uint64_t *full_bar,
                                               uint32_t parity)
{
    constexpr int PX = kFusedPX;
    const int tid = threadIdx.x;
    const int nx = p.nx, ny = p.ny;
    // staged rows / columns that outputs inside the array actually read: [0, rows_need) x [0, cols_need)
    const int rows_need = min(ny - gy0 + p.hi1, TYIN);
    const int cols_need = min(nx - gx0 + p.hi2, PX);
    // ... and the part of them that lies inside the array (TMA delivers it)
    const int r_lo = min(max(-gy0, 0), rows_need), r_hi = max(min(ny - gy0, rows_need), r_lo);
    const int c_lo = min(max(-gx0, 0), cols_need), c_hi = max(min(nx - gx0, cols_need), c_lo);
    const bool zero_ok = p.mode1 == B200_MODE_CONSTANT && p.mode2 == B200_MODE_CONSTANT && p.cval == 0.f;
    const bool any = r_lo > 0 || r_hi < rows_need || c_lo > 0 || c_hi < cols_need;
    const bool patch = TMA ? (any && !zero_ok) : true;
    if (patch && (TMA || any)) {
        // source row / column of every staged row / column (-1 = cval; identity inside the array)
        for (int i = tid; i < TYIN; i += NTHR) row_src[i] = ext_index32(gy0 + i, ny, p.mode1);
        for (int c = tid; c < PX; c += NTHR) col_src[c] = ext_index32(gx0 + c, nx, p.mode2);
    }
    __syncthreads();
    if constexpr (TMA) {
        if (tid < 32) mbar_wait(full_bar, parity);
        __syncthreads();
        if (patch) {
            // rows outside the array, full width
            const int nbad_r = r_lo + (rows_need - r_hi);
            for (int idx = tid; idx < nbad_r * PX; idx += NTHR) {
                const int k = idx / PX, c = idx - k * PX;
                if (c >= cols_need) continue;
                const int i = k < r_lo ? k : r_hi + (k - r_lo);
                const int ry = row_src[i], cx = col_src[c];
                tin[i * PX + c] = (ry < 0 || cx < 0) ? (TS)p.cval : plane[(int64_t)ry * nx + cx];
            }
            // columns outside the array on the rows inside it
            const int nbad_c = c_lo + (cols_need - c_hi);
            const int ngood_r = r_hi - r_lo;
            for (int idx = tid; idx < ngood_r * nbad_c; idx += NTHR) {
                const int r = idx / nbad_c, k = idx - r * nbad_c;
                const int c = k < c_lo ? k : c_hi + (k - c_lo);
                const int i = r_lo + r;
                const int cx = col_src[c];
                tin[i * PX + c] = cx < 0 ? (TS)p.cval : plane[(int64_t)(gy0 + i) * nx + cx];
            }
        }
    } else if constexpr (std::is_same<TS, float>::value) {
        // rows that are not 16-byte multiples (or an unaligned base): the same tile, staged with
        // plain 4-byte loads through the extension maps
        // plain 4-byte asynchronous copies (all of the tile in flight at once): straight from the rows
        // when the needed box lies inside the array, else through the extension maps (cval positions
        // are written directly)
        const int total = rows_need * PX;
        if (!any) {
            const int lane = tid & 31;
            for (int i = tid >> 5; i < rows_need; i += NTHR / 32) {
                const float *src = plane + (int64_t)(gy0 + i) * nx + gx0;
                float *dst = tin + i * PX;
#pragma unroll
                for (int c = lane; c < PX; c += 32) cp_async_f32(dst + c, c < cols_need ? src + c : plane, c < cols_need);
            }
        } else
        for (int idx = tid; idx < total; idx += NTHR) {
            const int i = idx / PX, c = idx - i * PX;
            const int ry = row_src[i], cx = c < cols_need ? col_src[c] : 0;
            if (c < cols_need && (ry < 0 || cx < 0)) {
                tin[idx] = p.cval;
            } else {
                const bool ok = c < cols_need;
                cp_async_f32(tin + idx, ok ? plane + (int64_t)ry * nx + cx : plane, ok);
            }
        }
        cp_async_wait_all();
    }
    __syncthreads();
}

// mode 'constant' along the contiguous axis: the unfused second pass reads cval beyond the ends of
// the STORED first pass, not the first pass of cval columns (they differ unless the taps of the
// first pass sum to exactly 1) -- overwrite those columns of the intermediate tile
template <int TYO, int NTHR>
__device__ __noinline__ void fused_fix_constant_columns(const Corr2dParams &p, float *mid, int gx0)
{
    constexpr int PX = kFusedPX, PI = kFusedPI;
    const int cols_need = min(p.nx - gx0 + p.hi2, PX);
    const int c_lo = min(max(-gx0, 0), cols_need), c_hi = max(min(p.nx - gx0, cols_need), c_lo);
    const int nbad_c = c_lo + (cols_need - c_hi);
    for (int idx = threadIdx.x; idx < TYO * nbad_c; idx += NTHR) {
        const int r = idx / nbad_c, k = idx - r * nbad_c;
        mid[r * PI + (k < c_lo ? k : c_hi + (k - c_lo))] = p.cval;
    }
    __syncthreads();
}

// 64-thread named barrier number 1 + idx with an IMMEDIATE barrier number (NPAIRS of them in use)
template <int NPAIRS>
__device__ __forceinline__ void pair_barrier(int idx)
{
    static_assert(NPAIRS >= 1 && NPAIRS <= 8, "warp pairs per CTA");
    if (NPAIRS == 1 || idx == 0) asm volatile("bar.sync 1, 64;" ::: "memory");
    else if (NPAIRS == 2 || idx == 1) asm volatile("bar.sync 2, 64;" ::: "memory");
    else if (NPAIRS == 3 || idx == 2) asm volatile("bar.sync 3, 64;" ::: "memory");
    else if (NPAIRS == 4 || idx == 3) asm volatile("bar.sync 4, 64;" ::: "memory");
    else if (NPAIRS == 5 || idx == 4) asm volatile("bar.sync 5, 64;" ::: "memory");
    else if (NPAIRS == 6 || idx == 5) asm volatile("bar.sync 6, 64;" ::: "memory");
    else if (NPAIRS == 7 || idx == 6) asm volatile("bar.sync 7, 64;" ::: "memory");
    else asm volatile("bar.sync 8, 64;" ::: "memory");
}

// One tile, from "the TMA load of the staged box has been issued on `full_bar`" to "its results are
// on their way to global memory": edge fill, both passes, staged stores.  `tin` is the staged box;
// the intermediate tile and the staged results alias it (ALIAS) or follow it.  `parity` is the phase
// of `full_bar` this tile's load completes.  Shared by the one-tile-per-CTA kernel and the persistent,
// double-buffered one.
template <int L1, int NTX, int TYO, int NWARPS, int RA, bool ALIAS, bool TMA, typename TS = float, bool SLIDE = false>
__device__ __forceinline__ void fused_tile(const Corr2dParams &p, float *tile, int *row_src, int *col_src,
                                           uint64_t *full_bar, uint32_t parity, int x0, int y0, int z)
{
    // BOX: the exact integer box filter (uniform_filter on u8 / u16).  The tile holds the storage values
    // as floats, window sums stay below 2^23 and are exact in any order, and each pass ends with the
    // quotient epilogue of the unfused kernels (box_quotient_bits) -- bit-exact by construction.
    constexpr bool BOX = !std::is_same<TS, float>::value;
    static_assert(!BOX || TMA, "the integer box pair is staged by TMA only");
    static_assert(!SLIDE || BOX, "sliding sums are exact for the integer box filter only");
    constexpr int PX = kFusedPX, PI = kFusedPI;
    constexpr int TYIN = TYO + L1 - 1;
    constexpr int NTHR = NWARPS * 32;
    constexpr int NCG = (PX - (NTX - 1)) / 16;   // 16-output column groups per tile, at most
    constexpr int NF = (16 + NTX + 3) / 4 * 4;   // row window of a thread: whole 16-byte blocks
    constexpr int ITEMS_A = (TYO / RA) * 2;
    static_assert(TYO % RA == 0 && TYO % 8 == 0, "tile rows");
    static_assert(!ALIAS || ITEMS_A == NWARPS, "aliased tiles need one phase-A item per warp");
    static_assert(16 * (NCG - 1) + NF <= PI, "row window inside the intermediate tile");
    // the staged box: floats, or the storage type of the integer box filter with the float intermediate
    // tile behind it (not aliased: phase A stores every row as soon as it is complete)
    TS *const tin = reinterpret_cast<TS *>(tile);
    float *mid = BOX ? reinterpret_cast<float *>(reinterpret_cast<unsigned char *>(tile) + box_raw_bytes<TS>(TYIN))
                     : (ALIAS ? tile : tile + TYIN * PX);
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int nx = p.nx, ny = p.ny;
    const int gx0 = x0 - p.lo2, gy0 = y0 - p.lo1;   // array position of staged element (0, 0)

    // a tile whose staged box lies inside the array needs nothing beyond the TMA load
    const bool edge = !TMA || gx0 < 0 || gy0 < 0 || gx0 + PX > nx || gy0 + TYIN > ny;
    if (edge) {
        fused_stage_edges<TYIN, NTHR, TMA, TS>(p, tin, row_src, col_src,
                                               static_cast<const TS *>(p.in) + (int64_t)z * ny * nx, gx0, gy0,
                                               full_bar, parity);
    } else {
        __syncthreads();
        // one warp polls the mbarrier; the others sleep at the CTA barrier
        if (tid < 32) mbar_wait(full_bar, parity);
        __syncthreads();
        B200_BOX_STAGE(1);
    }
    B200_BOX_STAGE(2);

    // ---- phase A: the pass along axis ndim-2, tin -> mid ----------------------
    for (int item = warp; item < ITEMS_A; item += NWARPS) {
        const int rg = item >> 1, half = item & 1;
        const int cv = (half * 32 + lane) * 4;
        if constexpr (BOX) {
            // integer box filter: rows are widened to float as they are loaded (box_load4), every output row is
            // stored as soon as it is complete -- the stored first pass, trunc(sum / size) exactly as
            // box_quotient_bits() rounds it, kept as a float (scalar FADD / FFMA as in the one-pass kernels)
            const TS *s = tin + rg * RA * PX + cv;
            float *d = mid + rg * RA * PI + cv;
            auto quot = [&](float sum) {
                return fmaf(SLIDE ? sum : sum + p.qc0, p.qinv, 12582912.f) - 12582912.f;   // (SLIDE: qc0 is in the sum)
            };
            auto put = [&](int j, float2 a, float2 b) {
                *reinterpret_cast<float4 *>(d + j * PI) = make_float4(quot(a.x), quot(a.y), quot(b.x), quot(b.y));
            };
            if constexpr (SLIDE) {
                // running column sums: row j + 1 = row j + (the row entering the window - the row leaving it);
                // the L1 rows of the window stay in registers, so every staged row is loaded once.  Rows are kept
                // as 2^23 + v (one byte permute per element): the first window is summed as r <- (r + row) - 2^23,
                // every step exact, and the differences of the sliding steps cancel the offset.  The sum carries
                // the quotient epilogue's pre-add qc0 = -(size - 1) / 2 (an integer for odd sizes) from the start.
                float4 v[RA + L1 - 1];
                const float2 k2 = make_float2(-8388608.f, -8388608.f);
                float2 r0 = make_float2(p.qc0, p.qc0), r1 = r0;
#pragma unroll
                for (int i = 0; i < L1; ++i) {
                    v[i] = box_load4<TS, true>(s + i * PX);
                    r0 = __fadd2_rn(__fadd2_rn(r0, make_float2(v[i].x, v[i].y)), k2);
                    r1 = __fadd2_rn(__fadd2_rn(r1, make_float2(v[i].z, v[i].w)), k2);
                }
                put(0, r0, r1);
#pragma unroll
                for (int j = 1; j < RA; ++j) {
                    v[j + L1 - 1] = box_load4<TS, true>(s + (j + L1 - 1) * PX);
                    const float4 vn = v[j + L1 - 1], vo = v[j - 1];
                    const float2 m1 = make_float2(-1.f, -1.f);
                    r0 = __fadd2_rn(r0, __ffma2_rn(make_float2(vo.x, vo.y), m1, make_float2(vn.x, vn.y)));
                    r1 = __fadd2_rn(r1, __ffma2_rn(make_float2(vo.z, vo.w), m1, make_float2(vn.z, vn.w)));
                    put(j, r0, r1);
                }
            } else {
                float2 acc[RA][2];
#pragma unroll
                for (int j = 0; j < RA; ++j) acc[j][0] = acc[j][1] = make_float2(0.f, 0.f);
#pragma unroll
                for (int i = 0; i < RA + L1 - 1; ++i) {
                    const float4 v = box_load4<TS>(s + i * PX);
#pragma unroll
                    for (int j = 0; j < RA; ++j) {
                        const int k = i - j;
                        if (k >= 0 && k < L1) {
                            const float2 w2 = make_float2(p.w1[k], p.w1[k]);
                            acc[j][0] = __ffma2_rn(make_float2(v.x, v.y), w2, acc[j][0]);
                            acc[j][1] = __ffma2_rn(make_float2(v.z, v.w), w2, acc[j][1]);
                        }
                    }
                }
#pragma unroll
                for (int j = 0; j < RA; ++j) put(j, acc[j][0], acc[j][1]);
            }
        } else {
            const float *s = tile + rg * RA * PX + cv;
            float2 acc[RA][2];
#pragma unroll
            for (int j = 0; j < RA; ++j) acc[j][0] = acc[j][1] = make_float2(0.f, 0.f);
#pragma unroll
            for (int i = 0; i < RA + L1 - 1; ++i) {
                const float4 v = *reinterpret_cast<const float4 *>(s + i * PX);
#pragma unroll
                for (int j = 0; j < RA; ++j) {
                    const int k = i - j;
                    if (k >= 0 && k < L1) {
                        const float wk = p.w1[k];
                        const float2 w2 = make_float2(wk, wk);
                        acc[j][0] = __ffma2_rn(make_float2(v.x, v.y), w2, acc[j][0]);
                        acc[j][1] = __ffma2_rn(make_float2(v.z, v.w), w2, acc[j][1]);
                    }
                }
            }
            if constexpr (ALIAS) __syncthreads();   // every warp has read its rows of tin
            float *d = mid + rg * RA * PI + cv;
#pragma unroll
            for (int j = 0; j < RA; ++j)
                *reinterpret_cast<float4 *>(d + j * PI) =
                    make_float4(acc[j][0].x, acc[j][0].y, acc[j][1].x, acc[j][1].y);
        }
    }
    // ---- phase B: the pass along the contiguous axis, mid -> out ---------------
    // Warp w owns rows [4w, 4w + 4) of the tile from here to the end: their intermediate values were
    // written by this warp and its partner (the two halves of row group w / 2), so ONE 64-thread
    // named barrier replaces a CTA-wide one, and nothing below synchronises beyond the warp.
    // Lane = (row, column group): the 8 lanes of a quarter warp sit on 4 rows x 2 adjacent column
    // groups, which the odd 16-byte pitch spreads over all 8 bank groups; a pass covers 8 column
    // groups, two passes the tile's 15.  Results go back into the rows IN PLACE (every window of a
    // pass is in registers before its first store; the second pass reads columns >= 128, the first
    // writes columns < 128) and leave the SM row by row through the bulk-copy engine (cp.async.bulk
    // shared -> global) instead of per-lane global stores, whose 32 half-sectors on 32 different
    // rows cost ~50 LSU wavefronts per instruction.  Epilogues that read `out`, and unaligned rows,
    // take a coalesced per-warp copy-out instead.
    static_assert(TYO == 4 * NWARPS && RA == 8 && ITEMS_A == NWARPS,
                  "a warp owns four rows of the tile in phase B, produced by itself and its partner warp");
    if (edge && p.mode2 == B200_MODE_CONSTANT && p.cval != 0.f) {
        __syncthreads();
        fused_fix_constant_columns<TYO, NTHR>(p, mid, gx0);
    } else {
        // (immediate barrier numbers: a register operand makes ptxas reserve all 16 hardware barriers
        // for the CTA, and an SM has 32 -- two resident CTAs instead of four)
        pair_barrier<NWARPS / 2>(warp >> 1);
    }
    B200_BOX_STAGE(3);
    const bool bulk = BOX || (TMA && p.epilogue == B200_EPI_STORE);
    const int tile_cols = min(nx - x0, p.ncg * 16);   // columns of this tile inside the array
    const int rw = 4 * warp;                           // first tile row of this warp
    float *const wrows = mid + rw * PI;
    const int rl = lane & 3;
    // Epilogues that read `out` (laplace / gradient-magnitude sums) fetch the previous values of this warp's
    // four rows -- 2 x 16 bytes per lane and row -- in one batch after the tap loops.  (Requesting them BEFORE
    // the tap loops hides their latency but keeps 32 more registers live: 104 registers, 2 CTAs per SM instead
    // of 3, and every fused launch -- C2 included -- ran 10-15 % slower on the B200.)
    constexpr bool EARLY_PREV = false;
    [[maybe_unused]] float4 pv[4][2];
    [[maybe_unused]] auto load_prev = [&]() {
        const bool reads_out = epilogue_reads_out(p.epilogue);
        const int nvec = tile_cols >> 2;
        const float *const dst0 = static_cast<const float *>(p.out) + ((int64_t)z * ny + y0 + rw) * nx + x0;
#pragma unroll
        for (int rr = 0; rr < 4; ++rr)
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int v = lane + 32 * u;
                pv[rr][u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (reads_out && y0 + rw + rr < ny && v < nvec)
                    pv[rr][u] = *reinterpret_cast<const float4 *>(dst0 + (int64_t)rr * nx + 4 * v);
            }
    };
    if constexpr (EARLY_PREV) {
        if (!bulk) load_prev();
    }
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
        const int cg = pass * 8 + 2 * (lane >> 3) + ((lane >> 2) & 1);
        const bool live = y0 + rw + rl < ny && cg * 16 < tile_cols;
        float *srow = wrows + rl * PI + cg * 16;
        float f[NF];
        float4 res[4];
        if (live) {
#pragma unroll
            for (int b = 0; b < NF / 4; ++b) {
                if constexpr (SLIDE) {
                    if (4 * b + 3 < NTX - L1) continue;     // in front of every window
                }
                const float4 v = *reinterpret_cast<const float4 *>(srow + 4 * b);
                f[4 * b] = v.x; f[4 * b + 1] = v.y; f[4 * b + 2] = v.z; f[4 * b + 3] = v.w;
            }
            if constexpr (SLIDE) {
                // running row sums over the window [i + LEAD, i + NTX) of output i
                constexpr int LEAD = NTX - L1;
                static_assert(LEAD >= 0 && LEAD < 16, "alignment zeros in front of the contiguous window");
                float rs = f[LEAD] + p.qc0;          // (the quotient epilogue's pre-add rides on the running sum)
#pragma unroll
                for (int t = LEAD + 1; t < NTX; ++t) rs += f[t];
                float r[16];
                r[0] = rs;
#pragma unroll
                for (int i = 1; i < 16; ++i) {
                    rs += f[i + NTX - 1] - f[i + LEAD - 1];
                    r[i] = rs;
                }
#pragma unroll
                for (int h = 0; h < 4; ++h) res[h] = make_float4(r[4 * h], r[4 * h + 1], r[4 * h + 2], r[4 * h + 3]);
            } else {
            float2 e[8], o[9];
#pragma unroll
            for (int i = 0; i < 8; ++i) e[i] = make_float2(0.f, 0.f);
#pragma unroll
            for (int i = 0; i < 9; ++i) o[i] = make_float2(0.f, 0.f);
#pragma unroll
            for (int tt = 0; tt < NTX; ++tt) {
                const float wk = p.w2[tt];
                const float2 w2 = make_float2(wk, wk);
                if (tt < 3 && wk == 0.f) continue;   // alignment zeros (as corr1d_contig_static_kernel)
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
#pragma unroll
            for (int h = 0; h < 4; ++h) {
                res[h].x = e[2 * h].x + o[2 * h].y;
                res[h].y = e[2 * h].y + o[2 * h + 1].x;
                res[h].z = e[2 * h + 1].x + o[2 * h + 1].y;
                res[h].w = e[2 * h + 1].y + o[2 * h + 2].x;
            }
            }
        }
        __syncwarp();      // every window of this pass has been read
        B200_BOX_STAGE(4);
        if (live) {
            if constexpr (BOX) {
                // quotients in the storage type, packed at the head of the row: output column c of the tile
                // sits at element c of a TS row (pass 0 fills the first 128 elements = at most 256 bytes,
                // below float column 128 where the windows of pass 1 start)
                TS *orow = reinterpret_cast<TS *>(wrows + rl * PI) + cg * 16;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const float s8[8] = {res[2 * h].x, res[2 * h].y, res[2 * h].z, res[2 * h].w,
                                         res[2 * h + 1].x, res[2 * h + 1].y, res[2 * h + 1].z, res[2 * h + 1].w};
                    if constexpr (SLIDE) {
                        uint32_t q8[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) q8[i] = __float_as_uint(fmaf(s8[i], p.qinv, 12582912.f));
                        pack_box_quotients<TS, 8>(orow + 8 * h, q8);
                    } else {
                        store_box_quotients<float, TS, 8>(orow + 8 * h, s8, p.qinv, p.qc0);
                    }
                }
            } else {
#pragma unroll
                for (int h = 0; h < 4; ++h) *reinterpret_cast<float4 *>(srow + 4 * h) = res[h];
            }
        }
    }
    B200_BOX_STAGE(5);
    if (bulk) fence_proxy_async();   // the stores above become visible to the bulk-copy engine
    __syncwarp();
    B200_BOX_STAGE(6);
    if (bulk) {
        if (lane < 4 && y0 + rw + lane < ny) {
            bulk_store_row(static_cast<TS *>(p.out) + ((int64_t)z * ny + y0 + rw + lane) * nx + x0, wrows + lane * PI,
                           tile_cols * (int)sizeof(TS));
            bulk_store_wait_read();    // the rows must stay intact until the engine has read them
        }
        return;
    }
    if constexpr (BOX) {
        return;
    } else if constexpr (TMA) {
        // 16-byte vectors, lanes along the row: coalesced read-modify-write with the epilogue.  All eight
        // loads of the previous values (4 rows x 2 vectors per lane) are issued before the first is used:
        // one round trip to L2 / HBM per tile instead of eight in a row.
        if constexpr (!EARLY_PREV) load_prev();
        const int nvec = tile_cols >> 2;
        float *const dst0 = static_cast<float *>(p.out) + ((int64_t)z * ny + y0 + rw) * nx + x0;
#pragma unroll
        for (int rr = 0; rr < 4; ++rr)
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int v = lane + 32 * u;
                if (y0 + rw + rr < ny && v < nvec) {
                    float4 r4 = *reinterpret_cast<const float4 *>(wrows + rr * PI + 4 * v);
                    r4.x = apply_epilogue<float>(p.epilogue, pv[rr][u].x, r4.x);
                    r4.y = apply_epilogue<float>(p.epilogue, pv[rr][u].y, r4.y);
                    r4.z = apply_epilogue<float>(p.epilogue, pv[rr][u].z, r4.z);
                    r4.w = apply_epilogue<float>(p.epilogue, pv[rr][u].w, r4.w);
                    *reinterpret_cast<float4 *>(dst0 + (int64_t)rr * nx + 4 * v) = r4;
                }
            }
        return;
    } else {
        const bool reads_out = epilogue_reads_out(p.epilogue);
        for (int rr = 0; rr < 4 && y0 + rw + rr < ny; ++rr) {
            const float *srcr = wrows + rr * PI;
            float *dst = static_cast<float *>(p.out) + ((int64_t)z * ny + y0 + rw + rr) * nx + x0;
            if (p.epilogue == B200_EPI_STORE) {
                for (int c = lane; c < tile_cols; c += 32) dst[c] = srcr[c];
            } else {
                for (int c = lane; c < tile_cols; c += 32)
                    dst[c] = apply_epilogue<float>(p.epilogue, reads_out ? dst[c] : 0.f, srcr[c]);
            }
        }
    }
}

template <int L1, int NTX, int TYO, int NWARPS, int RA, bool ALIAS, int MINB, bool TMA, typename TS = float,
          bool SLIDE = false>
__global__ void __launch_bounds__(NWARPS * 32, MINB)
corr2d_fused_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ Corr2dParams p)
{
    constexpr int PX = kFusedPX;
    constexpr int TYIN = TYO + L1 - 1;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *tin = aligned_smem<float>(smem_raw);
    __shared__ __align__(8) uint64_t full_bar;
    __shared__ int row_src[TYIN];
    __shared__ int col_src[PX];
    const int z = blockIdx.z;
    const int x0 = blockIdx.x * (p.ncg * 16), y0 = blockIdx.y * TYO;
    if constexpr (TMA) {
        if (threadIdx.x == 0) {
            mbar_init(&full_bar, 1);
            fence_barrier_init();
            mbar_arrive_expect_tx(&full_bar, (uint32_t)(TYIN * PX * (int)sizeof(TS)));
            tma_load_3d(tin, &tmap, &full_bar, x0 - p.lo2, y0 - p.lo1, z);
        }
    }
    fused_tile<L1, NTX, TYO, NWARPS, RA, ALIAS, TMA, TS, SLIDE>(p, tin, row_src, col_src, &full_bar, 0u, x0, y0, z);
}

// The same tiles on a PERSISTENT grid (as many CTAs as fit the device, each walking the tile list with
// the grid's stride) with the staged box double-buffered: the TMA load of a CTA's next-but-one tile is
// issued as soon as the buffer is free, i.e. a whole tile of arithmetic before it is needed, so no CTA
// ever sits at the barrier waiting for its data.  Two buffers per CTA halve the resident CTAs (2 x 8
// warps per SM), which the missing load latency more than pays for with short filters.
template <int L1, int NTX, int TYO, int NWARPS, int RA>
__global__ void __launch_bounds__(NWARPS * 32, 2)
corr2d_fused_persist_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ Corr2dParams p)
{
    constexpr int PX = kFusedPX;
    constexpr int TYIN = TYO + L1 - 1;
    constexpr int BUF = fused_buffer_floats(L1, TYO, NWARPS);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *base = aligned_smem<float>(smem_raw);
    __shared__ __align__(8) uint64_t full_bar[2];
    __shared__ int row_src[TYIN];
    __shared__ int col_src[PX];
    const int tid = threadIdx.x;
    const int G = gridDim.x;
    const int per_plane = p.x_tiles * p.y_tiles;
    const int total = p.n_tiles;
    auto issue = [&](int t, int b) {
        const int z = t / per_plane, q = t - z * per_plane;
        const int ty = q / p.x_tiles, tx = q - ty * p.x_tiles;
        mbar_arrive_expect_tx(&full_bar[b], (uint32_t)(TYIN * PX * (int)sizeof(float)));
        tma_load_3d(base + b * BUF, &tmap, &full_bar[b], tx * (p.ncg * 16) - p.lo2, ty * TYO - p.lo1, z);
    };
    int t = blockIdx.x;
    if (tid == 0) {
        mbar_init(&full_bar[0], 1);
        mbar_init(&full_bar[1], 1);
        fence_barrier_init();
        if (t < total) issue(t, 0);
        if (t + G < total) issue(t + G, 1);
    }
    for (int k = 0; t < total; ++k, t += G) {
        const int b = k & 1;
        const int z = t / per_plane, q = t - z * per_plane;
        const int ty = q / p.x_tiles, tx = q - ty * p.x_tiles;
        fused_tile<L1, NTX, TYO, NWARPS, RA, true, true>(p, base + b * BUF, row_src, col_src, &full_bar[b],
                                                          (uint32_t)(k >> 1) & 1u, tx * (p.ncg * 16), ty * TYO, z);
        // (every warp leaves fused_tile with its bulk stores read out of shared memory)
        __syncthreads();
        if (tid == 0 && t + 2 * G < total) {
            fence_proxy_async();     // the threads' accesses to this buffer precede the engine's writes
            issue(t + 2 * G, b);
        }
    }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
struct FusedCfg {
    int tyo, nwarps, ra, alias, minb;
};

template <int L1, int NTX, int TYO, int NWARPS, int RA, bool ALIAS, int MINB, bool TMA, typename TS = float,
          bool SLIDE = false>
static int launch_fused_cfg(const Corr2dArgs &a, const Corr2dParams &p0)
{
    constexpr int NCG = (kFusedPX - (NTX - 1)) / 16;
    constexpr int TXO = NCG * 16;
    constexpr int TYIN = TYO + L1 - 1;
    Corr2dParams p = p0;
    // as few tile columns as the widest tile allows, then the narrowest tile that still covers the row
    p.x_tiles = (int)((a.nx + TXO - 1) / TXO);
    p.ncg = (int)((a.nx + 16ll * p.x_tiles - 1) / (16ll * p.x_tiles));
    p.y_tiles = (int)((a.ny + TYO - 1) / TYO);
    if (a.outer > 65535 || p.y_tiles > 65535) return -1;   // grid = (x tiles, y tiles, planes)
    const size_t in_bytes = (size_t)TYIN * kFusedPX * sizeof(float);
    const size_t mid_bytes = (size_t)TYO * kFusedPI * sizeof(float);
    const size_t smem = std::is_same<TS, float>::value
                            ? (ALIAS ? (in_bytes > mid_bytes ? in_bytes : mid_bytes) : in_bytes + mid_bytes) + 128
                            : (size_t)box_raw_bytes<TS>(TYIN) + mid_bytes + 128;
    if (smem + 2048 > kSmemPerSm) return -1;
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    if constexpr (TMA) {
        uint64_t dims[3] = {(uint64_t)a.nx, (uint64_t)a.ny, (uint64_t)a.outer};
        uint64_t strides[2] = {(uint64_t)a.nx * sizeof(TS), (uint64_t)a.nx * sizeof(TS) * (uint64_t)a.ny};
        uint32_t bx[3] = {(uint32_t)kFusedPX, (uint32_t)TYIN, 1u};
        if (encode_tensor_map(&tmap, dtype_code<TS>(), (void *)a.in, 3, dims, strides, bx) != 0) return -1;
    }
    auto kern = corr2d_fused_kernel<L1, NTX, TYO, NWARPS, RA, ALIAS, MINB, TMA, TS, SLIDE>;
    B200_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemOptIn - 2048));
    kern<<<dim3((unsigned)p.x_tiles, (unsigned)p.y_tiles, (unsigned)a.outer), NWARPS * 32, smem, a.stream>>>(tmap, p);
    B200_CUDA_TRY(cudaGetLastError());
    count_launch();
    if constexpr (std::is_same<TS, float>::value) note_kernel(TMA ? "corr2d_fused_tma_f32" : "corr2d_fused_ldg_f32");
    else note_kernel(sizeof(TS) == 2 ? "uniform2d_fused_tma_u16" : "uniform2d_fused_tma_u8");
    return B200_OK;
}

template <int L1, int NTX>
static int launch_fused_persist(const Corr2dArgs &a, const Corr2dParams &p0)
{
    constexpr int TYO = 32, NWARPS = 8, RA = 8;
    constexpr int NCG = (kFusedPX - (NTX - 1)) / 16;
    constexpr int TXO = NCG * 16;
    constexpr int TYIN = TYO + L1 - 1;
    constexpr size_t smem = 2 * (size_t)fused_buffer_floats(L1, TYO, NWARPS) * sizeof(float) + 128;
    Corr2dParams p = p0;
    p.x_tiles = (int)((a.nx + TXO - 1) / TXO);
    p.ncg = (int)((a.nx + 16ll * p.x_tiles - 1) / (16ll * p.x_tiles));
    p.y_tiles = (int)((a.ny + TYO - 1) / TYO);
    const int64_t n_tiles = (int64_t)p.x_tiles * p.y_tiles * a.outer;
    if (n_tiles >= (1ll << 31)) return -1;
    p.n_tiles = (int)n_tiles;
    CUtensorMap tmap;
    uint64_t dims[3] = {(uint64_t)a.nx, (uint64_t)a.ny, (uint64_t)a.outer};
    uint64_t strides[2] = {(uint64_t)a.nx * 4, (uint64_t)a.nx * 4 * (uint64_t)a.ny};
    uint32_t bx[3] = {(uint32_t)kFusedPX, (uint32_t)TYIN, 1u};
    if (encode_tensor_map(&tmap, B200_F32, (void *)a.in, 3, dims, strides, bx) != 0) return -1;
    auto kern = corr2d_fused_persist_kernel<L1, NTX, TYO, NWARPS, RA>;
    B200_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemOptIn - 2048));
    const int64_t resident = 2ll * device_sm_count();
    const unsigned grid = (unsigned)(n_tiles < resident ? n_tiles : resident);
    kern<<<grid, NWARPS * 32, smem, a.stream>>>(tmap, p);
    B200_CUDA_TRY(cudaGetLastError());
    count_launch();
    note_kernel("corr2d_fused_tma_f32");
    return B200_OK;
}

// tile geometry: 32 output rows x up to 240 columns, 8 warps, 8 rows per lane in phase A, the
// intermediate tile aliased onto the input tile (B200 sweep: tools/sweep_fused.py with a
// -DB200_FUSED_SWEEP build, profiles/r2_sweep_fused.log)
template <int L1, int NTX, bool TMA>
static int launch_fused_taps(const Corr2dArgs &a, const Corr2dParams &p)
{
    const int variant = env_int("B200_FUSED_VARIANT", 0);
#ifdef B200_FUSED_SWEEP
    if constexpr (TMA && (L1 == 17 || L1 == 25 || L1 == 33))
    switch (variant) {
    case 1: return launch_fused_cfg<L1, NTX, 32, 8, 8, false, 2, TMA>(a, p);
    case 2: return launch_fused_cfg<L1, NTX, 32, 8, 8, true, 2, TMA>(a, p);
    case 3: return launch_fused_cfg<L1, NTX, 32, 8, 8, true, 4, TMA>(a, p);
    case 4: return launch_fused_cfg<L1, NTX, 16, 4, 8, true, 6, TMA>(a, p);
    case 6: return launch_fused_cfg<L1, NTX, 64, 16, 8, true, 1, TMA>(a, p);
    case 7: return launch_fused_cfg<L1, NTX, 24, 6, 8, true, 5, TMA>(a, p);   // 24 rows x 6 warps, 5 CTAs / SM
    case 8: return launch_fused_cfg<L1, NTX, 24, 6, 8, true, 4, TMA>(a, p);
    default: break;
    }
#else
    (void)variant;
#endif
    // persistent double-buffered form where two CTAs with two staged boxes each fit an SM (<= 23 taps)
    if constexpr (TMA && 2 * (2 * fused_buffer_floats(L1, 32, 8) * 4 + 128 + 2560) <= (int)kSmemPerSm) {
        if (env_int("B200_FUSED_PERSIST", 0) && a.epilogue == B200_EPI_STORE) return launch_fused_persist<L1, NTX>(a, p);
    }
    // 64 registers for every filter length: 4 CTAs of 8 warps per SM below 25 taps; from 25 taps on shared memory
    // admits 3 CTAs either way, and the 64-register code still runs 0-10 % faster than the 72-register one
    // (B200 sweeps at 1024^3 and 512 x 2048^2, profiles/r2_sweep_fused_minb_*.log)
    return launch_fused_cfg<L1, NTX, 32, 8, 8, true, 4, TMA>(a, p);
}

// (L1, NTX) pairs with a compiled kernel: the two passes of an isotropic odd
// filter with origin 0 (every gaussian: radius lw, L = 2 lw + 1, lead = alignment
// zeros in front of the contiguous taps), lw = 1 ... 23
#define B200_FUSED_PAIRS(X) \
    X(3, 6) X(5, 7) X(7, 8) X(9, 9) X(11, 14) X(13, 15) X(15, 16) X(17, 17) X(19, 22) X(21, 23) X(23, 24) \
    X(25, 25) X(27, 30) X(29, 31) X(31, 32) X(33, 33) X(41, 41)

static bool fused_pair_compiled(int l1, int ntx)
{
#define X(A, B) if (l1 == A && ntx == B) return true;
    B200_FUSED_PAIRS(X)
#undef X
    return false;
}

// `align` = elements per 16 bytes: the staged box starts lo2_al columns in front of the tile, and TMA wants
// the byte offset of that column to be a multiple of 16 (a float tile: 4 elements; uint16: 8; uint8: 16 --
// a misaligned inner coordinate faults with "illegal instruction")
static void fused_geometry(int nw2, int origin2, int *lo2_al, int *lead, int *ntx, int align = 4)
{
    const int lo2 = nw2 / 2 + origin2;
    *lo2_al = (lo2 + align - 1) / align * align;
    *lead = *lo2_al - lo2;
    *ntx = *lead + nw2;
}

int corr2d_fused_eligible(int dtype, int64_t ny, int64_t nx, int nw1, int origin1, int nw2, int origin2)
{
    if (dtype != B200_F32 || nw1 < 1 || nw2 < 1) return 0;
    if (origin1 < -(nw1 / 2) || origin1 > (nw1 - 1) / 2 || origin2 < -(nw2 / 2) || origin2 > (nw2 - 1) / 2) return 0;
    if (ny < 1 || nx < 1 || ny >= (1ll << 31) || nx >= (1ll << 31)) return 0;
    int lo2_al, lead, ntx;
    fused_geometry(nw2, origin2, &lo2_al, &lead, &ntx);
    return fused_pair_compiled(nw1, ntx) ? 1 : 0;
}

int launch_corr2d_fused(const Corr2dArgs &a)
{
    if (!corr2d_fused_eligible(a.dtype, a.ny, a.nx, a.nw1, a.origin1, a.nw2, a.origin2)) return -1;
    if (a.outer >= (1ll << 31)) return -1;
    int lo2_al, lead, ntx;
    fused_geometry(a.nw2, a.origin2, &lo2_al, &lead, &ntx);

    Corr2dParams p;
    memset(&p, 0, sizeof(p));
    p.in = a.in; p.out = a.out;
    p.outer = (int)a.outer; p.ny = (int)a.ny; p.nx = (int)a.nx;
    p.lo1 = a.nw1 / 2 + a.origin1; p.mode1 = a.mode1; p.hi1 = a.nw1 - 1 - p.lo1;
    p.lo2 = lo2_al; p.mode2 = a.mode2; p.hi2 = ntx - 1 - lo2_al;
    p.epilogue = a.epilogue;
    p.wide_store = (((a.nx * 4) % 32) == 0 && ((uintptr_t)a.out & 31) == 0) ? 1 : 0;
    p.cval = (float)a.cval;
    for (int i = 0; i < a.nw1; ++i) p.w1[i] = (float)a.w1[i];
    for (int i = 0; i < a.nw2; ++i) p.w2[lead + i] = (float)a.w2[i];

    // TMA needs 16-byte rows and bases; anything else is staged with plain loads
    const bool tma = ((a.nx * 4) % 16 == 0) && !((uintptr_t)a.in & 15) && !((uintptr_t)a.out & 15) &&
                     !(a.flags & B200_FLAG_NO_TMA);
#define X(A, B)                                                                              \
    if (a.nw1 == A && ntx == B)                                                              \
        return tma ? launch_fused_taps<A, B, true>(a, p) : launch_fused_taps<A, B, false>(a, p);
    B200_FUSED_PAIRS(X)
#undef X
    return -1;
}

// ---------------------------------------------------------------------------
// the exact integer box filter (uniform_filter on u8 / u16) along the last two axes in one launch
// ---------------------------------------------------------------------------
// isotropic odd sizes with origin 0 (the (L1, NTX) geometry of the gaussian pairs); the running-sum
// form costs the same for every size
#define B200_BOX_PAIRS_U16(X) X(3, 10) X(5, 11) X(7, 12) X(9, 13) X(11, 14) X(13, 15) X(15, 16) X(17, 17)
#define B200_BOX_PAIRS_U8(X) X(3, 18) X(5, 19) X(7, 20) X(9, 21) X(11, 22) X(13, 23) X(15, 24) X(17, 25)

int uniform2d_fused_eligible(int dtype, int64_t ny, int64_t nx, int size1, int origin1, int mode1, int size2, int origin2,
                             int mode2, double cval)
{
    if (dtype != B200_U16 && dtype != B200_U8) return 0;
    if (size1 != size2 || origin1 != 0 || origin2 != 0 || !(size1 & 1)) return 0;
    if (ny < 1 || nx < 1 || ny >= (1ll << 31) || nx >= (1ll << 31)) return 0;
    if ((nx * (dtype == B200_U16 ? 2 : 1)) % 16) return 0;          // TMA rows and bulk row stores
    if (!uniform_box_ok(dtype, size1, mode1, cval) || !uniform_box_ok(dtype, size2, mode2, cval)) return 0;
    int lo2_al, lead, ntx;
    fused_geometry(size2, 0, &lo2_al, &lead, &ntx, dtype == B200_U16 ? 8 : 16);
#define X(A, B) if (size1 == A && ntx == B) return 1;
    if (dtype == B200_U16) {
        B200_BOX_PAIRS_U16(X)
    } else {
        B200_BOX_PAIRS_U8(X)
    }
#undef X
    return 0;
}

static int launch_uniform2d_u16(const Corr2dArgs &a, const Corr2dParams &p, int ntx)
{
    // B200_BOX2D_SLIDE=0: the direct form (one add per tap) of the 7-wide filter, kept for the A/B
    if (env_int("B200_BOX2D_SLIDE", 1) == 0 && a.nw1 == 7 && ntx == 12)
        return launch_fused_cfg<7, 12, 32, 8, 8, true, 4, true, unsigned short, false>(a, p);
#define X(A, B) \
    if (a.nw1 == A && ntx == B) return launch_fused_cfg<A, B, 32, 8, 8, true, 4, true, unsigned short, true>(a, p);
    B200_BOX_PAIRS_U16(X)
#undef X
    return -1;
}

static int launch_uniform2d_u8(const Corr2dArgs &a, const Corr2dParams &p, int ntx)
{
#define X(A, B) \
    if (a.nw1 == A && ntx == B) return launch_fused_cfg<A, B, 32, 8, 8, true, 4, true, unsigned char, true>(a, p);
    B200_BOX_PAIRS_U8(X)
#undef X
    return -1;
}

int launch_uniform2d_fused(const Uniform2dArgs &u)
{
    if (!uniform2d_fused_eligible(u.dtype, u.ny, u.nx, u.size1, u.origin1, u.mode1, u.size2, u.origin2, u.mode2, u.cval))
        return -1;
    if (u.outer >= (1ll << 31) || ((uintptr_t)u.in & 15) || ((uintptr_t)u.out & 15) || (u.flags & B200_FLAG_NO_TMA))
        return -1;
    int lo2_al, lead, ntx;
    fused_geometry(u.size2, u.origin2, &lo2_al, &lead, &ntx, u.dtype == B200_U16 ? 8 : 16);
    Corr2dArgs a;
    memset(&a, 0, sizeof(a));
    a.in = u.in; a.out = u.out; a.dtype = u.dtype; a.outer = u.outer; a.ny = u.ny; a.nx = u.nx;
    a.nw1 = u.size1; a.nw2 = u.size2; a.mode1 = u.mode1; a.mode2 = u.mode2; a.cval = u.cval;
    a.epilogue = B200_EPI_STORE; a.flags = u.flags; a.stream = u.stream;
    Corr2dParams p;
    memset(&p, 0, sizeof(p));
    p.in = u.in; p.out = u.out;
    p.outer = (int)u.outer; p.ny = (int)u.ny; p.nx = (int)u.nx;
    p.lo1 = u.size1 / 2; p.mode1 = u.mode1; p.hi1 = u.size1 - 1 - p.lo1;
    p.lo2 = lo2_al; p.mode2 = u.mode2; p.hi2 = ntx - 1 - lo2_al;
    p.epilogue = B200_EPI_STORE;
    p.cval = (float)u.cval;
    p.qinv = 1.0f / (float)u.size1;
    p.qc0 = 0.5f - 0.5f * (float)u.size1;
#ifdef B200_BOX_DEBUG
    p.dbg = env_int("B200_BOX_DBG", 0);
#endif
    for (int i = 0; i < u.size1; ++i) p.w1[i] = 1.f;
    for (int i = 0; i < u.size2; ++i) p.w2[lead + i] = 1.f;
    return u.dtype == B200_U16 ? launch_uniform2d_u16(a, p, ntx) : launch_uniform2d_u8(a, p, ntx);
}

}  // namespace b200
