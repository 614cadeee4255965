// corr1d_minmax.cu -- minimum / maximum filter along the CONTIGUOUS axis on the
// static contiguous-axis engine of corr1d_contig.cu (float32, uint8 / uint16
// storage): same tile, lane mapping and stores, every tap is 16 FMNMX.
#include "corr1d_contig.cuh"

namespace b200 {

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

// ---------------------------------------------------------------------------
// entry point (corr1d_tma.cuh)
// ---------------------------------------------------------------------------
int tma_contig_minmax(const Corr1dArgs &a, bool is_max)
{
    switch (a.dtype) {
    case B200_F32: return is_max ? launch_minmax_contig<float, true>(a) : launch_minmax_contig<float, false>(a);
    case B200_U16: return is_max ? launch_minmax_contig<uint16_t, true>(a) : launch_minmax_contig<uint16_t, false>(a);
    case B200_U8: return is_max ? launch_minmax_contig<uint8_t, true>(a) : launch_minmax_contig<uint8_t, false>(a);
    default: return -1;
    }
}

}  // namespace b200
