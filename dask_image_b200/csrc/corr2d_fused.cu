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

struct Corr2dParams {
    const float *in;
    float *out;
    int64_t outer, ny, nx;
    int lo1, mode1;       // rows read above an output row: nw1/2 + origin1
    int lo2, mode2;       // staged column 0 is array column x0 - lo2 (lo2 = reach in front, rounded up to 4)
    int hi1, hi2;         // rows / columns read beyond an output
    int epilogue, wide_store;
    int y_tiles, x_tiles;
    float cval;
    float w1[kFusedMaxTaps];   // taps along axis ndim-2
    float w2[kFusedMaxTaps];   // taps along axis ndim-1, alignment zeros in front
};

template <int L1, int NTX, int TYO, int NWARPS, int RA, bool ALIAS, int MINB, bool TMA>
__global__ void __launch_bounds__(NWARPS * 32, MINB)
corr2d_fused_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ Corr2dParams p)
{
    constexpr int PX = kFusedPX, PI = kFusedPI;
    constexpr int TYIN = TYO + L1 - 1;
    constexpr int NTHR = NWARPS * 32;
    constexpr int NCG = (PX - (NTX - 1)) / 16;   // 16-output column groups per tile
    constexpr int TXO = NCG * 16;
    constexpr int NF = (16 + NTX + 3) / 4 * 4;   // row window of a thread: whole 16-byte blocks
    constexpr int ITEMS_A = (TYO / RA) * 2;
    constexpr int ITEMS_B = TYO * NCG;
    static_assert(TYO % RA == 0 && TYO % 8 == 0, "tile rows");
    static_assert(!ALIAS || ITEMS_A == NWARPS, "aliased tiles need one phase-A item per warp");
    static_assert(16 * (NCG - 1) + NF <= PI, "row window inside the intermediate tile");

    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *tin = aligned_smem<float>(smem_raw);
    float *mid = ALIAS ? tin : tin + TYIN * PX;
    __shared__ __align__(8) uint64_t full_bar;
    __shared__ int row_src[TYIN];
    __shared__ int col_src[PX];

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;

    const int64_t t = blockIdx.x;
    const int txi = (int)(t % p.x_tiles);
    const int64_t q = t / p.x_tiles;
    const int tyi = (int)(q % p.y_tiles);
    const int64_t z = q / p.y_tiles;
    const int64_t x0 = (int64_t)txi * TXO, y0 = (int64_t)tyi * TYO;
    const int64_t gx0 = x0 - p.lo2, gy0 = y0 - p.lo1;   // array position of staged element (0, 0)

    if constexpr (TMA) {
        if (tid == 0) {
            mbar_init(&full_bar, 1);
            fence_barrier_init();
            mbar_arrive_expect_tx(&full_bar, (uint32_t)(TYIN * PX * (int)sizeof(float)));
            tma_load_3d(tin, &tmap, &full_bar, (int)gx0, (int)gy0, (int)z);
        }
    }

    // Staged rows / columns that are actually read by outputs inside the array
    // (a tile on the far edges may hang over): [0, rows_need) x [0, cols_need)
    const int rows_need = (int)(p.ny - gy0 + p.hi1 < TYIN ? p.ny - gy0 + p.hi1 : TYIN);
    const int cols_need = (int)(p.nx - gx0 + p.hi2 < PX ? p.nx - gx0 + p.hi2 : PX);
    // ... and the part of that which lies inside the array (TMA delivers it; the rest is zero-filled)
    const int r_lo = (int)(gy0 < 0 ? (-gy0 < rows_need ? -gy0 : rows_need) : 0);
    const int r_hi = (int)(p.ny - gy0 < rows_need ? (p.ny - gy0 > r_lo ? p.ny - gy0 : r_lo) : rows_need);
    const int c_lo = (int)(gx0 < 0 ? (-gx0 < cols_need ? -gx0 : cols_need) : 0);
    const int c_hi = (int)(p.nx - gx0 < cols_need ? (p.nx - gx0 > c_lo ? p.nx - gx0 : c_lo) : cols_need);
    const bool edge = r_lo > 0 || r_hi < rows_need || c_lo > 0 || c_hi < cols_need;
    const bool zero_ok = p.mode1 == B200_MODE_CONSTANT && p.mode2 == B200_MODE_CONSTANT && p.cval == 0.f;
    const bool patch = TMA ? (edge && !zero_ok) : true;
    const float *plane = p.in + z * p.ny * p.nx;

    if (patch && (edge || !TMA)) {
        // source row / column of every staged row / column under the extension maps (-1 = cval);
        // identity inside the array
        for (int i = tid; i < TYIN; i += NTHR) row_src[i] = (int)ext_index(gy0 + i, p.ny, p.mode1);
        for (int c = tid; c < PX; c += NTHR) col_src[c] = (int)ext_index(gx0 + c, p.nx, p.mode2);
    }
    __syncthreads();

    if constexpr (TMA) {
        // one warp polls the mbarrier; the others sleep at the CTA barrier
        if (tid < 32) mbar_wait(&full_bar, 0);
        __syncthreads();
        if (patch) {
            // rows outside the array, full width
            const int nbad_r = r_lo + (rows_need - r_hi);
            for (int idx = tid; idx < nbad_r * PX; idx += NTHR) {
                const int k = idx / PX, c = idx - k * PX;
                if (c >= cols_need) continue;
                const int i = k < r_lo ? k : r_hi + (k - r_lo);
                const int ry = row_src[i], cx = col_src[c];
                tin[i * PX + c] = (ry < 0 || cx < 0) ? p.cval : plane[(int64_t)ry * p.nx + cx];
            }
            // columns outside the array on the rows inside it
            const int nbad_c = c_lo + (cols_need - c_hi);
            const int ngood_r = r_hi - r_lo;
            for (int idx = tid; idx < ngood_r * nbad_c; idx += NTHR) {
                const int r = idx / nbad_c, k = idx - r * nbad_c;
                const int c = k < c_lo ? k : c_hi + (k - c_lo);
                const int i = r_lo + r;
                const int cx = col_src[c];
                tin[i * PX + c] = cx < 0 ? p.cval : plane[(gy0 + i) * p.nx + cx];
            }
            __syncthreads();
        }
    } else {
        // rows that are not 16-byte multiples (or an unaligned base): the same tile, staged with
        // plain 4-byte loads through the extension maps
        const int total = rows_need * PX;
#pragma unroll 4
        for (int idx = tid; idx < total; idx += NTHR) {
            const int i = idx / PX, c = idx - i * PX;
            float v = 0.f;
            if (c < cols_need) {
                const int ry = row_src[i], cx = col_src[c];
                v = (ry < 0 || cx < 0) ? p.cval : plane[(int64_t)ry * p.nx + cx];
            }
            tin[idx] = v;
        }
        __syncthreads();
    }

    // ---- phase A: the pass along axis ndim-2, tin -> mid ----------------------
    for (int item = warp; item < ITEMS_A; item += NWARPS) {
        const int rg = item >> 1, half = item & 1;
        const int cv = (half * 32 + lane) * 4;
        const float *s = tin + rg * RA * PX + cv;
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
            *reinterpret_cast<float4 *>(d + j * PI) = make_float4(acc[j][0].x, acc[j][0].y, acc[j][1].x, acc[j][1].y);
    }
    __syncthreads();
    // mode 'constant' along the contiguous axis: the unfused second pass reads cval beyond the
    // ends of the STORED first pass, not the first pass of cval columns (they differ unless the
    // taps of the first pass sum to exactly 1) -- overwrite those columns of the intermediate
    if (p.mode2 == B200_MODE_CONSTANT && p.cval != 0.f && (c_lo > 0 || c_hi < cols_need)) {
        const int nbad_c = c_lo + (cols_need - c_hi);
        for (int idx = tid; idx < TYO * nbad_c; idx += NTHR) {
            const int r = idx / nbad_c, k = idx - r * nbad_c;
            mid[r * PI + (k < c_lo ? k : c_hi + (k - c_lo))] = p.cval;
        }
        __syncthreads();
    }

    // ---- phase B: the pass along the contiguous axis, mid -> out ---------------
    // item = (column group, row): the 8 lanes of a quarter warp sit on 8 rows of one
    // column group, which the odd 16-byte pitch spreads over all bank groups
    for (int item = tid; item < ITEMS_B; item += NTHR) {
        const int cg = item / TYO, r = item - cg * TYO;
        const float *srow = mid + r * PI + cg * 16;
        float f[NF];
#pragma unroll
        for (int b = 0; b < NF / 4; ++b) {
            const float4 v = *reinterpret_cast<const float4 *>(srow + 4 * b);
            f[4 * b] = v.x; f[4 * b + 1] = v.y; f[4 * b + 2] = v.z; f[4 * b + 3] = v.w;
        }
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
        const int64_t row = y0 + r;
        const int64_t x = x0 + cg * 16;
        if (row < p.ny && x < p.nx) {
            float *dst0 = p.out + (z * p.ny + row) * p.nx + x;
            float res[16];
#pragma unroll
            for (int j = 0; j < 16; ++j)
                res[j] = ((j & 1) ? e[j >> 1].y : e[j >> 1].x) + ((j & 1) ? o[(j + 1) >> 1].x : o[j >> 1].y);
            if (p.wide_store && x + 16 <= p.nx) {
                if (p.epilogue != B200_EPI_STORE) {
                    const bool reads_out = epilogue_reads_out(p.epilogue);
                    float prev[16];
                    if (reads_out) {
                        ld_global_256(dst0, reinterpret_cast<float(&)[8]>(prev[0]));
                        ld_global_256(dst0 + 8, reinterpret_cast<float(&)[8]>(prev[8]));
                    }
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        res[j] = apply_epilogue<float>(p.epilogue, reads_out ? prev[j] : 0.f, res[j]);
                }
                st_global_256(dst0, reinterpret_cast<const float(&)[8]>(res[0]));
                st_global_256(dst0 + 8, reinterpret_cast<const float(&)[8]>(res[8]));
            } else if constexpr (TMA) {
#pragma unroll
                for (int h = 0; h < 4; ++h) {
                    if (x + h * 4 < p.nx) {   // rows are whole multiples of 16 bytes
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
            } else {
                const bool reads_out = epilogue_reads_out(p.epilogue);
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    if (x + j < p.nx) {
                        float v = res[j];
                        if (p.epilogue != B200_EPI_STORE)
                            v = apply_epilogue<float>(p.epilogue, reads_out ? dst0[j] : 0.f, v);
                        dst0[j] = v;
                    }
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
struct FusedCfg {
    int tyo, nwarps, ra, alias, minb;
};

template <int L1, int NTX, int TYO, int NWARPS, int RA, bool ALIAS, int MINB, bool TMA>
static int launch_fused_cfg(const Corr2dArgs &a, const Corr2dParams &p0)
{
    constexpr int NCG = (kFusedPX - (NTX - 1)) / 16;
    constexpr int TXO = NCG * 16;
    constexpr int TYIN = TYO + L1 - 1;
    Corr2dParams p = p0;
    p.x_tiles = (int)((a.nx + TXO - 1) / TXO);
    p.y_tiles = (int)((a.ny + TYO - 1) / TYO);
    const int64_t num_tiles = (int64_t)p.x_tiles * p.y_tiles * a.outer;
    if (num_tiles >= (1ll << 31)) return -1;
    const size_t in_bytes = (size_t)TYIN * kFusedPX * sizeof(float);
    const size_t mid_bytes = (size_t)TYO * kFusedPI * sizeof(float);
    const size_t smem = (ALIAS ? (in_bytes > mid_bytes ? in_bytes : mid_bytes) : in_bytes + mid_bytes) + 128;
    if (smem + 2048 > kSmemPerSm) return -1;
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    if constexpr (TMA) {
        uint64_t dims[3] = {(uint64_t)a.nx, (uint64_t)a.ny, (uint64_t)a.outer};
        uint64_t strides[2] = {(uint64_t)a.nx * 4, (uint64_t)a.nx * 4 * (uint64_t)a.ny};
        uint32_t bx[3] = {(uint32_t)kFusedPX, (uint32_t)TYIN, 1u};
        if (encode_tensor_map(&tmap, B200_F32, (void *)a.in, 3, dims, strides, bx) != 0) return -1;
    }
    auto kern = corr2d_fused_kernel<L1, NTX, TYO, NWARPS, RA, ALIAS, MINB, TMA>;
    B200_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemOptIn - 2048));
    kern<<<(unsigned)num_tiles, NWARPS * 32, smem, a.stream>>>(tmap, p);
    B200_CUDA_TRY(cudaGetLastError());
    count_launch();
    note_kernel(TMA ? "corr2d_fused_tma_f32" : "corr2d_fused_ldg_f32");
    return B200_OK;
}

// tile geometry: 32 output rows, 8 warps, 8 rows per lane in phase A, the
// intermediate tile aliased onto the input tile (B200 sweep: tools/sweep_fused.py)
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
    case 5: return launch_fused_cfg<L1, NTX, 64, 8, 8, false, 1, TMA>(a, p);
    case 6: return launch_fused_cfg<L1, NTX, 64, 16, 8, true, 1, TMA>(a, p);
    case 7: return launch_fused_cfg<L1, NTX, 32, 4, 16, true, 4, TMA>(a, p);
    default: break;
    }
#else
    (void)variant;
#endif
    return launch_fused_cfg<L1, NTX, 32, 8, 8, true, 3, TMA>(a, p);
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

static void fused_geometry(int nw2, int origin2, int *lo2_al, int *lead, int *ntx)
{
    const int lo2 = nw2 / 2 + origin2;
    *lo2_al = (lo2 + 3) / 4 * 4;
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
    p.in = (const float *)a.in; p.out = (float *)a.out;
    p.outer = a.outer; p.ny = a.ny; p.nx = a.nx;
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

}  // namespace b200
