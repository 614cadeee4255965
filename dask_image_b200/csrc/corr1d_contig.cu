// corr1d_contig.cu -- TMA-tiled 1-D pass along the CONTIGUOUS axis (sm_100a).
//
//  * corr1d_contig_static_kernel -- float32 (or uint8/uint16 storage: the exact
//    box filter) and at most 48 taps: the tap count is a template parameter and
//    the tap loop is straight-line packed FFMA2.
//  * corr1d_contig_tma_kernel -- the run-time-loop form (float64, longer filters).
//  (corr1d_minmax.cu holds the static engine with FMNMX taps.)
//
// View [rows][n]; a CTA stages a TR x P box with one cp.async.bulk.tensor.2d
// whose pitch P is an odd multiple of 16 bytes, so the lanes of a quarter warp
// hit 8 distinct 16-byte bank groups.  Out-of-range columns are zero-filled by
// TMA; CTAs on an array edge fetch exactly those elements through the closed-form
// extension map while the tile is in flight.
#include "corr1d_contig.cuh"

namespace b200 {

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
// alignment zeros in front of a box of `size` taps with origin 0 (the TMA box starts on a 16-byte boundary of
// the storage type: EB elements)
constexpr int box_lead(int size, int EB) { return (size / 2 + EB - 1) / EB * EB - size / 2; }
constexpr int box_taps(int size, int EB) { return (box_lead(size, EB) + size + EB - 1) / EB * EB; }

// BS > 0 (integer box filter of BS taps, origin 0): the window slides along the thread's 16 outputs --
// BS adds for the first, one add and one subtract for every further one -- instead of BS taps each.
template <typename TS, int NT, bool LDG = false, int BS = 0>
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

    contig_stage_tile<TS, 8, LDG>(tmap, tile, &full_bar, p.in, g0, row0, p.n, nrows, P, TR, p.mode, p.cval);

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
        float slide[R];
        if constexpr (BS > 0) {
            // sliding window: sums of integers below 2^23 are exact in fp32 in any order
            static_assert(kBox, "the sliding window is the exact integer box filter");
            constexpr int BL = box_lead(BS, EB);
            static_assert(BL + BS <= NT && R - 1 + BL + BS <= NF, "window inside the staged row");
            float sacc = 0.f;
#pragma unroll
            for (int tt = 0; tt < BS; ++tt) sacc += f[BL + tt];
            slide[0] = sacc;
#pragma unroll
            for (int j = 1; j < R; ++j) {
                sacc = (sacc + f[j - 1 + BL + BS]) - f[j - 1 + BL];
                slide[j] = sacc;
            }
        } else {
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
        }
        const int64_t row = row0 + rg * 8 + r_in;
        if (patch_live && row < nrows && x < p.n) {
            TS *dst0 = out_thr + (int64_t)(rg * 8) * p.n;
            float res[R];
#pragma unroll
            for (int j = 0; j < R; ++j) {
                if constexpr (BS > 0)
                    res[j] = slide[j];
                else
                    res[j] = ((j & 1) ? e[j >> 1].y : e[j >> 1].x) + ((j & 1) ? o[(j + 1) >> 1].x : o[j >> 1].y);
            }
            if constexpr (LDG) {
                // unaligned rows: element stores, the row may end anywhere
                const bool reads_out = epilogue_reads_out(p.epilogue);
#pragma unroll
                for (int j = 0; j < R; ++j) {
                    if (x + j < p.n) {
                        if constexpr (kBox) {
                            dst0[j] = (TS)box_quotient_bits<float>(res[j], p.qinv, p.qc0);
                        } else {
                            dst0[j] = p.epilogue == B200_EPI_STORE
                                          ? res[j]
                                          : apply_epilogue<float>(p.epilogue, reads_out ? dst0[j] : 0.f, res[j]);
                        }
                    }
                }
            } else if constexpr (kBox) {
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

template <typename TS, int NT, bool LDG = false, int BS = 0>
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
    memset(&tmap, 0, sizeof(tmap));
    if constexpr (!LDG) {
        uint64_t dims[2] = {(uint64_t)g.n, (uint64_t)g.outer};
        uint64_t strides[1] = {(uint64_t)g.n * es};
        uint32_t bx[2] = {(uint32_t)pitch, (uint32_t)tr};
        if (encode_tensor_map(&tmap, dtype_code<TS>(), (void *)a.in, 2, dims, strides, bx) != 0) return -1;
    }

    Corr1dTmaParams<float, TS> p;
    p.in = (const TS *)a.in; p.out = (TS *)a.out;
    p.outer = g.outer; p.n = g.n; p.inner = 1;
    p.pad_lo = 0; p.pad_hi = 0;
    p.nw = NT; p.lo = lo_al; p.mode = a.mode; p.epilogue = a.epilogue;
    p.tile_rows = tr; p.smem_rows = tr; p.nb = 1; p.pitch = pitch; p.shift = 0; p.ncg = ncg;
    p.wide_store = (!LDG && ((g.n * es) % 32) == 0 && ((uintptr_t)a.out & 31) == 0) ? 1 : 0;
    p.n_tiles = (int)r_tiles; p.c_tiles = (int)c_tiles;
    p.cval = (TS)a.cval;
    p.qinv = box.qinv; p.qc0 = box.qc0;
    for (int i = 0; i < NT; ++i) p.w[i] = 0.f;
    for (int i = 0; i < a.nw; ++i) p.w[lead + i] = (float)a.w[i];

    auto kern = corr1d_contig_static_kernel<TS, NT, LDG, BS>;
    const size_t smem = stage_bytes + 128;
    B200_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemOptIn));
    kern<<<(unsigned)num_tiles, threads == 128 ? 128 : 256, smem, a.stream>>>(tmap, p);
    B200_CUDA_TRY(cudaGetLastError());
    count_launch();
    note_kernel(kernel_name<float, TS>(false, LDG));
    return B200_OK;
}

// float32 rows with at most kStaticTaps taps (alignment zeros in front
// included): the compile-time tap engine.  Returns -2 when the filter is longer.
constexpr int kStaticTaps = 48;
static int launch_contig_static(const Corr1dArgs &a)
{
    const LineGeom &g = a.g;
    if (a.pad_lo || a.pad_hi) return -1;
    if (g.n >= (1ll << 31) || g.outer >= (1ll << 31)) return -1;
    // rows / bases that are not 16-byte multiples: the same tiles staged with plain loads
    const bool ldg = (g.n * 4) % 16 != 0 || ((uintptr_t)a.in & 15) || ((uintptr_t)a.out & 15);
    const int lo = a.nw / 2 + a.origin;
    const int lo_al = (lo + 3) / 4 * 4;
    const int lead = lo_al - lo;
    const int nt = lead + a.nw;
    if (nt > kStaticTaps) return ldg ? -1 : -2;
    const BoxInfo none;
    switch (nt) {
#define B200_C1_CASE(K)                                                                  \
    case K:                                                                              \
        return ldg ? launch_contig_static_cfg<float, K, true>(a, none, lo_al, lead)      \
                   : launch_contig_static_cfg<float, K, false>(a, none, lo_al, lead);
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
    if (g.n >= (1ll << 31) || g.outer >= (1ll << 31)) return -1;
    const bool ldg = (g.n * es) % 16 != 0 || ((uintptr_t)a.in & 15) || ((uintptr_t)a.out & 15);
    const int lo = a.nw / 2 + a.origin;
    const int lo_al = (lo + U16B - 1) / U16B * U16B;
    const int lead = lo_al - lo;
    if (!ldg && a.origin == 0 && env_int("B200_BOX_SLIDE", 1)) {
        // common windows: the sliding-window instantiation (window size and alignment zeros at compile time)
#define B200_SLIDE_CASE(S)                                                                              \
    case S:                                                                                             \
        return launch_contig_static_cfg<TS, box_taps(S, U16B), false, S>(a, box, lo_al, lead);
        switch (a.nw) {
        B200_SLIDE_CASE(2) B200_SLIDE_CASE(3) B200_SLIDE_CASE(4) B200_SLIDE_CASE(5) B200_SLIDE_CASE(6)
        B200_SLIDE_CASE(7) B200_SLIDE_CASE(8) B200_SLIDE_CASE(9) B200_SLIDE_CASE(10) B200_SLIDE_CASE(11)
        B200_SLIDE_CASE(12) B200_SLIDE_CASE(13) B200_SLIDE_CASE(14) B200_SLIDE_CASE(15) B200_SLIDE_CASE(16)
        B200_SLIDE_CASE(17) B200_SLIDE_CASE(19) B200_SLIDE_CASE(21) B200_SLIDE_CASE(25) B200_SLIDE_CASE(31)
        default: break;
        }
#undef B200_SLIDE_CASE
    }
    if (ldg || env_int("B200_TMA_STATIC", 1)) {
        // compile-time tap engine: taps padded to whole 16-byte blocks of storage
        const int la = (lead + a.nw + U16B - 1) / U16B * U16B;
#define B200_BOX_CASE(K)                                                              \
    return ldg ? launch_contig_static_cfg<TS, K, true>(a, box, lo_al, lead)           \
               : launch_contig_static_cfg<TS, K, false>(a, box, lo_al, lead);
        switch (la) {
        case 8: if constexpr (U16B == 8) { B200_BOX_CASE(8) } break;
        case 16: B200_BOX_CASE(16)
        case 24: if constexpr (U16B == 8) { B200_BOX_CASE(24) } break;
        case 32: B200_BOX_CASE(32)
        case 40: if constexpr (U16B == 8) { B200_BOX_CASE(40) } break;
        case 48: B200_BOX_CASE(48)
        default: break;
        }
#undef B200_BOX_CASE
    }
    if (ldg) return -1;
    int L = a.nw + lead;
    L = (L + V - 1) / V * V;
    if (L > kMaxTapsTma) return -1;
    return launch_contig_cfg<float, TS, 0, 0>(a, box, lo_al, L);
}


// ---------------------------------------------------------------------------
// entry points (corr1d_tma.cuh)
// ---------------------------------------------------------------------------
int tma_contig_corr(const Corr1dArgs &a)
{
    if (a.dtype == B200_F32) {
        // (which contiguous-axis engine runs depends on the tap count only)
        const int rc = env_int("B200_TMA_STATIC", 1) ? launch_contig_static(a) : -2;
        return rc == -2 ? launch_contig<float>(a) : rc;
    }
    if (a.dtype == B200_F64) return launch_contig<double>(a);
    return -1;
}

int tma_contig_box(const Corr1dArgs &a, const BoxInfo &box)
{
    if (a.dtype == B200_U16) return launch_contig_box<uint16_t>(a, box);
    if (a.dtype == B200_U8) return launch_contig_box<uint8_t>(a, box);
    return -1;
}

}  // namespace b200
