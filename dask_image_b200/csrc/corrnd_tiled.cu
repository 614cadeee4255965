// corrnd_tiled.cu -- shared-memory-tiled direct N-D correlate (sm_100a) for
// non-separable weights: the float path behind convolve / correlate
// (replaces _nd_image.correlate, scipy/ndimage/_filters.py:1305).
//
// The array is viewed as [Z][Y][X] (X contiguous; 2-D arrays have Z = 1).  A CTA
// stages an input tile of (TZ + Kz - 1) x (TY + Ky - 1) rows of P elements with
// ONE cp.async.bulk.tensor.3d (P is an odd multiple of 16 bytes: conflict-free
// LDS.128), TMA zero-fills what lies outside the array, edge CTAs patch exactly
// those elements through the closed-form extension maps.  A thread then owns
// R = 2V consecutive outputs of one row and walks the Kz x Ky weight rows; each
// row is the contiguous-axis tap engine of the 1-D kernel (tile_common.cuh:
// three rotating 16-byte register blocks, packed FFMA2 on aligned pairs,
// weights on the uniform datapath straight out of the parameter bank):
// per V taps 1 LDS.128 + R*V FMAs.  The tap order per voxel is fixed
// (kz, ky, kx), independent of tile geometry and slab partition.
//
// Bound: 2*taps flop / 8 bytes per voxel -- FP32 issue for anything beyond a few
// dozen taps (9x9x9: 729 FMA per voxel), HBM for small stencils.
// fp32 (fp64 for float64 arrays) accumulation; B200_FLAG_EXACT selects the
// scipy-order fp64 kernel in generic.cu instead.  Weights with |w| <= DBL_EPSILON
// are treated as zero, as the reference skips them.
#include "tile_common.cuh"
#include <vector>
#include <cmath>

namespace b200 {

constexpr int kNdTapBytes = 24 * 1024;   // weights live in the kernel parameter bank

template <typename T>
struct CorrNdTiledParams {
    const T *in;
    T *out;
    int64_t X, Y, Z;          // extents of the planes this call produces
    int64_t pad_lo, pad_hi;   // resident planes before / after along the slowest axis
    int pad_dim;              // which tile dimension the pads belong to: 2 = z, 1 = y (2-D arrays)
    int Kx_al, Ky, Kz;        // taps per weight row (incl. SH alignment zeros), rows, planes
    int lox_al, loy, loz;     // tile origin = first output - these
    int TY, TZ, ncg;          // output tile: (ncg * 4R) x TY x TZ
    int P, BY, BZ;            // input tile: pitch x rows x planes
    int cx, cy;               // tiles along x and y (z = the rest of the grid)
    int mode, wide_store;
    T cval;
    // neighbour planes read in place (multi-GPU slab, corrnd_static_peer_kernel): the n_peer_lo planes that
    // precede plane 0 and the n_peer_hi planes that follow plane Z - 1, in a peer GPU's memory; null = none
    const T *peer_lo, *peer_hi;
    int n_peer_lo, n_peer_hi;
    T w[kNdTapBytes / sizeof(T)];
};

// Tensor maps of the peer-halo form (3-D arrays, slabs along z): the first z tile takes its top `loz`
// planes from `lo` (the previous rank's last planes) and the rest from a shallower local box; the last
// z tile takes the planes behind the slab from `hi`.  Every other tile is unchanged.
struct NdPeerMaps {
    CUtensorMap first;   // local slab, box depth BZ - loz
    CUtensorMap last;    // local slab, box depth k_last
    CUtensorMap lo;      // the planes that precede the slab
    CUtensorMap hi;      // the planes that follow it
    int has_lo, has_hi;
    int lo_planes, k_last, ncz;
};

// Stage the CTA's input tile: one TMA 3-D box (zero fill outside the resident
// range), then -- on CTAs that touch an edge -- patch exactly the out-of-range
// elements through the extension maps.  Ends with the tile visible to all.
template <typename T, typename PARAMS, bool PEER = false>
__device__ __forceinline__ void nd_stage_tile(const PARAMS &p, const CUtensorMap &tmap, T *tile, uint64_t *full_bar,
                                              int P, int BY, int BZ, int64_t g0x, int64_t g0y, int64_t g0z,
                                              int64_t ylo, int64_t yhi, int64_t zlo, int64_t zhi,
                                              const NdPeerMaps *hm = nullptr, int64_t tcz = 0)
{
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        mbar_init(full_bar, 1);
        fence_barrier_init();
        mbar_arrive_expect_tx(full_bar, (uint32_t)((size_t)P * BY * BZ * sizeof(T)));
        if constexpr (PEER) {
            // (the slab's tensor maps have no pad planes in front: z coordinates are slab planes)
            if (tcz == 0 && hm->has_lo) {
                tma_load_3d(tile, &hm->lo, full_bar, (int)g0x, (int)g0y, 0);
                tma_load_3d(tile + (size_t)hm->lo_planes * BY * P, &hm->first, full_bar, (int)g0x, (int)g0y, 0);
            } else if (tcz == hm->ncz - 1 && hm->has_hi) {
                tma_load_3d(tile, &hm->last, full_bar, (int)g0x, (int)g0y, (int)g0z);
                tma_load_3d(tile + (size_t)hm->k_last * BY * P, &hm->hi, full_bar, (int)g0x, (int)g0y, 0);
            } else {
                tma_load_3d(tile, &tmap, full_bar, (int)g0x, (int)g0y, (int)g0z);
            }
        } else {
            tma_load_3d(tile, &tmap, full_bar, (int)g0x, (int)(g0y - ylo), (int)(g0z - zlo));
        }
    }
    // plane `sz` of the resident range: local, or one of the neighbours' planes
    auto plane_base = [&](int64_t sz) -> const T * {
        if constexpr (PEER) {
            if (sz < 0) return p.peer_lo + (sz + p.n_peer_lo) * p.Y * p.X;
            if (sz >= p.Z) return p.peer_hi + (sz - p.Z) * p.Y * p.X;
        }
        return p.in + sz * p.Y * p.X;
    };
    __syncthreads();
    if (tid < 32) mbar_wait(full_bar, 0);
    __syncthreads();

    const bool x_edge = g0x < 0 || g0x + P > p.X;
    const bool yz_edge = g0y < ylo || g0y + BY > yhi || g0z < zlo || g0z + BZ > zhi;
    if ((x_edge || yz_edge) && !(p.mode == B200_MODE_CONSTANT && p.cval == T(0))) {
        for (int rr = warp; rr < BY * BZ; rr += 8) {
            const int zz = rr / BY, r = rr - zz * BY;
            const int64_t gy = g0y + r, gz = g0z + zz;
            const bool row_in = gy >= ylo && gy < yhi && gz >= zlo && gz < zhi;
            if (row_in && !x_edge) continue;
            int64_t sy = gy, sz = gz;
            bool row_cval = false;
            if (!row_in) {
                sy = p.pad_dim == 1 ? ext_index_padded(gy, p.Y, p.mode, p.pad_lo, p.pad_hi)
                                    : ext_index_padded(gy, p.Y, p.mode, 0, 0);
                sz = p.pad_dim == 2 ? ext_index_padded(gz, p.Z, p.mode, p.pad_lo, p.pad_hi)
                                    : ext_index_padded(gz, p.Z, p.mode, 0, 0);
                row_cval = sy == B200_USE_CVAL || sz == B200_USE_CVAL;
            }
            const T *srow = row_cval ? nullptr : plane_base(sz) + sy * p.X;
            T *trow = tile + (size_t)rr * P;
            for (int c = lane; c < P; c += 32) {
                const int64_t gx = g0x + c;
                const bool x_in = gx >= 0 && gx < p.X;
                if (row_in && x_in) continue;
                T v = p.cval;
                if (srow) {
                    const int64_t sx = ext_index(gx, p.X, p.mode);
                    if (sx >= 0) v = srow[sx];
                }
                trow[c] = v;
            }
        }
        __syncthreads();
    }

}

template <typename T, int SH, int REM>
__global__ void __launch_bounds__(256, 2)
corrnd_tiled_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CorrNdTiledParams<T> p)
{
    constexpr int V = 16 / (int)sizeof(T);
    constexpr int R = 2 * V;
    constexpr int PC = 4 * R;
    using VecT = Vec<T, V>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    T *tile = aligned_smem<T>(smem_raw);
    __shared__ __align__(8) uint64_t full_bar;

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int P = p.P, BY = p.BY, BZ = p.BZ;
    const int SX = p.ncg * PC;

    int64_t t = blockIdx.x;
    const int tcx = (int)(t % p.cx);
    t /= p.cx;
    const int tcy = (int)(t % p.cy);
    const int64_t tcz = t / p.cy;
    const int64_t x0t = (int64_t)tcx * SX, y0t = (int64_t)tcy * p.TY, z0t = tcz * p.TZ;
    const int64_t g0x = x0t - p.lox_al, g0y = y0t - p.loy, g0z = z0t - p.loz;
    // resident range of y and z (pads belong to the slowest axis of the array)
    const int64_t ylo = p.pad_dim == 1 ? -p.pad_lo : 0, yhi = p.Y + (p.pad_dim == 1 ? p.pad_hi : 0);
    const int64_t zlo = p.pad_dim == 2 ? -p.pad_lo : 0, zhi = p.Z + (p.pad_dim == 2 ? p.pad_hi : 0);

    nd_stage_tile<T>(p, tmap, tile, &full_bar, P, BY, BZ, g0x, g0y, g0z, ylo, yhi, zlo, zhi);

    const int nrg = p.TY / 8;
    const int per_plane = nrg * p.ncg;
    const int npatch = per_plane * p.TZ;
    const int patch_iters = (npatch + 7) / 8;   // uniform trip count, see corr1d_contig_tma_kernel
    const int rows_w = p.Ky * p.Kz;
    for (int qi = 0; qi < patch_iters; ++qi) {
        const int q_raw = qi * 8 + warp;
        const bool patch_live = q_raw < npatch;
        const int q = patch_live ? q_raw : npatch - 1;
        const int zz = q / per_plane;
        const int qq = q - zz * per_plane;
        const int rg = qq % nrg, cg = qq / nrg;
        const int r = rg * 8 + (lane >> 2);
        const int xl = cg * PC + (lane & 3) * R;
        const T *sbase = tile + ((size_t)zz * BY + r) * P + xl;
        ContigAcc<T, V> acc;
        acc.zero();
        const T *wrow = p.w;
        int ky = 0, kz = 0;
        for (int wr = 0; wr < rows_w; ++wr, wrow += p.Kx_al) {
            row_taps<T, T, V, SH, REM>(acc, sbase + ((size_t)kz * BY + ky) * P, wrow, p.Kx_al);
            if (++ky == p.Ky) { ky = 0; ++kz; }
        }
        const int64_t x = x0t + xl, y = y0t + r, z = z0t + zz;
        if (patch_live && y < p.Y && z < p.Z && x < p.X) {
            T *dst0 = p.out + ((z * p.Y + y) * p.X + x);
            if (p.wide_store && x + R <= p.X) {
                T res[R];
#pragma unroll
                for (int j = 0; j < R; ++j) res[j] = acc.get(j);
                st_global_256(dst0, res);
            } else {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    if (x + h * V < p.X) {
                        VecT res;
#pragma unroll
                        for (int v = 0; v < V; ++v) res.v[v] = acc.get(h * V + v);
                        *reinterpret_cast<VecT *>(dst0 + h * V) = res;
                    }
                }
            }
        }
    }
}

// ===========================================================================
// float32 specialisation with the weight-row length fixed at compile time
// (KXA = taps per row incl. alignment zeros, 1..16): the row engine is
// straight-line code -- the thread's 16 + KXA input floats in registers, every
// tap a run of packed FFMA2 with its weight read through the uniform datapath
// at a static offset from the row's base -- so a row costs its FMAs, its 16-byte
// shared loads and a handful of loop instructions.  A thread owns 16 consecutive
// outputs (odd taps pair outputs (-1,0)...(15,16): 9 FFMA2 instead of 8); a warp
// covers 64 x 8 outputs, lanes of a quarter warp sit on 8 different rows so the
// odd 16-byte pitch makes every LDS.128 conflict-free.
// ===========================================================================
template <int KXA, bool SKIPZ, bool PEER>
__device__ __forceinline__ void corrnd_static_body(const CUtensorMap &tmap, const CorrNdTiledParams<float> &p,
                                                   const NdPeerMaps *hm)
{
    constexpr int R = 16;
    constexpr int NF = (R + KXA + 3) / 4 * 4;   // floats of a row a thread touches
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *tile = aligned_smem<float>(smem_raw);
    __shared__ __align__(8) uint64_t full_bar;

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int P = p.P, BY = p.BY, BZ = p.BZ;
    constexpr int SX = 4 * R;

    int64_t t = blockIdx.x;
    const int tcx = (int)(t % p.cx);
    t /= p.cx;
    const int tcy = (int)(t % p.cy);
    const int64_t tcz = t / p.cy;
    const int64_t x0t = (int64_t)tcx * SX, y0t = (int64_t)tcy * p.TY, z0t = tcz * p.TZ;
    const int64_t g0x = x0t - p.lox_al, g0y = y0t - p.loy, g0z = z0t - p.loz;
    const int64_t ylo = p.pad_dim == 1 ? -p.pad_lo : 0, yhi = p.Y + (p.pad_dim == 1 ? p.pad_hi : 0);
    const int64_t zlo = p.pad_dim == 2 ? -p.pad_lo : 0, zhi = p.Z + (p.pad_dim == 2 ? p.pad_hi : 0);

    nd_stage_tile<float, CorrNdTiledParams<float>, PEER>(p, tmap, tile, &full_bar, P, BY, BZ, g0x, g0y, g0z, ylo, yhi,
                                                          zlo, zhi, hm, tcz);

    // patches: (plane zz, row group yg); TY / 8 is a power of two
    const int nyg = p.TY >> 3;
    const int nyg_shift = 31 - __clz(nyg);
    const int npatch = nyg * p.TZ;
    const int patch_iters = (npatch + 7) >> 3;
    const int r_in = lane & 7, xq = lane >> 3;
    const int Ky = p.Ky, Kz = p.Kz;
    for (int qi = 0; qi < patch_iters; ++qi) {
        const int q_raw = qi * 8 + warp;
        const bool patch_live = q_raw < npatch;
        const int q = patch_live ? q_raw : npatch - 1;
        const int zz = q >> nyg_shift, yg = q & (nyg - 1);
        const int r = yg * 8 + r_in;
        const int xl = xq * R;
        const float *sz = tile + ((size_t)zz * BY + r) * P + xl;
        float2 e[8], o[9];
#pragma unroll
        for (int i = 0; i < 8; ++i) e[i] = make_float2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 9; ++i) o[i] = make_float2(0.f, 0.f);
        const float *wrow = p.w;
        for (int kz = 0; kz < Kz; ++kz, sz += (size_t)BY * P) {
            const float *srow = sz;
            for (int ky = 0; ky < Ky; ++ky, srow += P, wrow += KXA) {
                float f[NF];
#pragma unroll
                for (int b = 0; b < NF / 4; ++b) {
                    const float4 v = *reinterpret_cast<const float4 *>(srow + 4 * b);
                    f[4 * b] = v.x; f[4 * b + 1] = v.y; f[4 * b + 2] = v.z; f[4 * b + 3] = v.w;
                }
#pragma unroll
                for (int tt = 0; tt < KXA; ++tt) {
                    const float wk = wrow[tt];
                    const float2 w2 = make_float2(wk, wk);
                    // Zero taps are SKIPPED, as the reference iterates over the non-zero
                    // footprint only (|w| > DBL_EPSILON): 0 * NaN / 0 * Inf must not reach the
                    // sum.  The alignment padding in front of a row (tt < 3) is always tested;
                    // kernels with genuine zero weights run the SKIPZ instantiation, which tests
                    // every tap (warp-uniform branch), dense kernels keep the untested form.
                    if ((SKIPZ || tt < 3) && wk == 0.f) continue;
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
        }
        const int64_t x = x0t + xl, y = y0t + r, z = z0t + zz;
        if (patch_live && y < p.Y && z < p.Z && x < p.X) {
            float *dst0 = p.out + ((z * p.Y + y) * p.X + x);
            float res[R];
#pragma unroll
            for (int j = 0; j < R; ++j)
                res[j] = ((j & 1) ? e[j >> 1].y : e[j >> 1].x) + ((j & 1) ? o[(j + 1) >> 1].x : o[j >> 1].y);
            if (p.wide_store && x + R <= p.X) {
                st_global_256(dst0, reinterpret_cast<const float(&)[8]>(res[0]));
                st_global_256(dst0 + 8, reinterpret_cast<const float(&)[8]>(res[8]));
            } else {
#pragma unroll
                for (int h = 0; h < 4; ++h) {
                    if (x + h * 4 < p.X)
                        *reinterpret_cast<float4 *>(dst0 + h * 4) =
                            make_float4(res[4 * h], res[4 * h + 1], res[4 * h + 2], res[4 * h + 3]);
                }
            }
        }
    }
}

template <int KXA, bool SKIPZ>
__global__ void __launch_bounds__(256, 2)
corrnd_static_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CorrNdTiledParams<float> p)
{
    corrnd_static_body<KXA, SKIPZ, false>(tmap, p, nullptr);
}

// the same kernel for a multi-GPU slab whose neighbour planes are read where they live (NdPeerMaps)
template <int KXA, bool SKIPZ>
__global__ void __launch_bounds__(256, 2)
corrnd_static_peer_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ NdPeerMaps hm,
                          const __grid_constant__ CorrNdTiledParams<float> p)
{
    corrnd_static_body<KXA, SKIPZ, true>(tmap, p, &hm);
}

template <typename T, int SH, int REM>
static int launch_cfg(const CorrNdArgs &a, int lox_al, int Kx_al)
{
    constexpr int V = 16 / (int)sizeof(T);
    constexpr int R = 2 * V, PC = 4 * R;
    const int es = (int)sizeof(T);
    const int nd = a.ndim;
    const int64_t X = a.shape[nd - 1], Y = a.shape[nd - 2], Z = nd == 3 ? a.shape[0] : 1;
    const int Kx = (int)a.wshape[nd - 1], Ky = (int)a.wshape[nd - 2], Kz = nd == 3 ? (int)a.wshape[0] : 1;
    const int loy = Ky / 2 + (int)a.origins[nd - 2];
    const int loz = nd == 3 ? Kz / 2 + (int)a.origins[0] : 0;
    const int SH_ = lox_al - (Kx / 2 + (int)a.origins[nd - 1]);

    auto pitch_for = [&](int cand) {
        int pt = cand * PC + Kx_al + 2 * V - 1;
        pt = (pt + V - 1) / V;
        if (!(pt & 1)) ++pt;
        return pt * V;
    };
    int ncg = 2;
    while (ncg > 1 && pitch_for(ncg) > 256) --ncg;
    const int P = pitch_for(ncg);
    if (P > 256) return -1;
    // tile depth / height: as many output planes (rows for 2-D arrays) as keep
    // two CTAs resident per SM; geometry never changes a voxel's tap order
    const size_t budget = (kSmemPerSm - 4096) / 2 - 256;
    int TY = 8, TZ = 1;
    auto bytes_for = [&](int ty, int tz) { return (size_t)P * (ty + Ky - 1) * (tz + Kz - 1) * es; };
    if (bytes_for(TY, TZ) > budget) return -1;
    if (nd == 3) {
        while (TZ < 8 && bytes_for(TY, TZ * 2) <= budget) TZ *= 2;
    } else {
        while (TY < 64 && bytes_for(TY * 2, 1) <= budget) TY *= 2;
    }
    const int BY = TY + Ky - 1, BZ = TZ + Kz - 1;
    if (BY > 256 || BZ > 256) return -1;
    const int SX = ncg * PC;
    const int64_t cx = (X + SX - 1) / SX, cy = (Y + TY - 1) / TY, cz = (Z + TZ - 1) / TZ;
    const int64_t num_tiles = cx * cy * cz;
    if (num_tiles >= (1ll << 31) || cx >= (1ll << 30) || cy >= (1ll << 30)) return -1;

    const int pad_dim = nd == 3 ? 2 : 1;
    const int64_t Ytot = pad_dim == 1 ? a.pad_lo + Y + a.pad_hi : Y;
    const int64_t Ztot = pad_dim == 2 ? a.pad_lo + Z + a.pad_hi : Z;
    const int64_t plane = pad_dim == 2 ? Y * X : X;
    const char *base = (const char *)a.in - a.pad_lo * plane * es;
    if (X >= (1ll << 31) || Ytot >= (1ll << 31) || Ztot >= (1ll << 31)) return -1;

    CUtensorMap tmap;
    uint64_t dims[3] = {(uint64_t)X, (uint64_t)Ytot, (uint64_t)Ztot};
    uint64_t strides[2] = {(uint64_t)X * es, (uint64_t)X * es * (uint64_t)Ytot};
    uint32_t bx[3] = {(uint32_t)P, (uint32_t)BY, (uint32_t)BZ};
    if (encode_tensor_map(&tmap, a.dtype, (void *)base, 3, dims, strides, bx) != 0) return -1;

    static_assert(sizeof(CorrNdTiledParams<T>) < 31 * 1024, "parameter bank");
    CorrNdTiledParams<T> p;
    p.in = (const T *)a.in; p.out = (T *)a.out;
    p.X = X; p.Y = Y; p.Z = Z;
    p.pad_lo = a.pad_lo; p.pad_hi = a.pad_hi; p.pad_dim = pad_dim;
    p.Kx_al = Kx_al; p.Ky = Ky; p.Kz = Kz;
    p.lox_al = lox_al; p.loy = loy; p.loz = loz;
    p.TY = TY; p.TZ = TZ; p.ncg = ncg; p.P = P; p.BY = BY; p.BZ = BZ;
    p.cx = (int)cx; p.cy = (int)cy;
    p.mode = a.mode; p.cval = (T)a.cval;
    p.peer_lo = p.peer_hi = nullptr; p.n_peer_lo = p.n_peer_hi = 0;
    p.wide_store = (((X * es) % 32) == 0 && ((uintptr_t)a.out & 31) == 0) ? 1 : 0;
    const int rows = Ky * Kz;
    for (int i = 0; i < rows * Kx_al; ++i) p.w[i] = T(0);
    for (int rw = 0; rw < rows; ++rw)
        for (int kx = 0; kx < Kx; ++kx) {
            const double w = a.w[(size_t)rw * Kx + kx];
            p.w[rw * Kx_al + SH_ + kx] = std::fabs(w) > DBL_EPSILON ? (T)w : T(0);
        }

    auto kern = corrnd_tiled_kernel<T, SH, REM>;
    const size_t smem = (size_t)P * BY * BZ * es + 128;
    B200_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemOptIn));
    kern<<<(unsigned)num_tiles, 256, smem, a.stream>>>(tmap, p);
    B200_CUDA_TRY(cudaGetLastError());
    count_launch();
    note_kernel(sizeof(T) == 4 ? "corrnd_tiled_f32" : "corrnd_tiled_f64");
    return B200_OK;
}

// Tile geometry of the static engine for a weight row of `kxa` taps (alignment zeros included): the
// deepest output tile 64 x TY x TZ that keeps two CTAs per SM resident.  false = no geometry.
static bool nd_static_tile(int nd, int kxa, int Ky, int Kz, int *P_out, int *TY_out, int *TZ_out)
{
    constexpr int R = 16, SX = 4 * R;
    const int NF = (R + kxa + 3) / 4 * 4;
    int pu = (SX - R + NF + 3) / 4;
    if (!(pu & 1)) ++pu;
    const int P = pu * 4;
    const size_t budget = (kSmemPerSm - 4096) / 2 - 256;
    static const int cand3[][2] = {{8, 8}, {16, 4}, {32, 2}, {8, 4}, {16, 2}, {8, 2}, {16, 1}, {8, 1}};
    static const int cand2[][2] = {{64, 1}, {32, 1}, {16, 1}, {8, 1}};
    const int (*cand)[2] = nd == 3 ? cand3 : cand2;
    const int ncand = nd == 3 ? 8 : 4;
    for (int i = 0; i < ncand; ++i) {
        const size_t bytes = (size_t)P * (cand[i][0] + Ky - 1) * (cand[i][1] + Kz - 1) * 4;
        if (bytes <= budget && cand[i][0] + Ky - 1 <= 256 && cand[i][1] + Kz - 1 <= 256) {
            *P_out = P; *TY_out = cand[i][0]; *TZ_out = cand[i][1];
            return true;
        }
    }
    return false;
}

// Peer-halo form: neighbour planes may enter the first and the last z tile only.
static bool nd_peer_tiles_ok(int64_t Z, int TZ, int loz, int hiz)
{
    const int64_t ncz = (Z + TZ - 1) / TZ;
    const int64_t last_planes = Z - (ncz - 1) * TZ;
    return ncz >= 2 && TZ >= loz && last_planes >= hiz;
}

// host-only: can launch_corrnd_peer take a slab of this shape with these weights?  (every rank evaluates it
// for every slab size, so that all of them choose the same path)
int corrnd_peer_eligible(int dtype, int ndim, const int64_t *shape, const int64_t *wshape, const int64_t *origins)
{
    if (dtype != B200_F32 || ndim != 3) return 0;
    const int64_t X = shape[2], Y = shape[1], Z = shape[0];
    if ((X * 4) % 16 != 0 || X >= (1ll << 31) || Y >= (1ll << 31) || Z >= (1ll << 31)) return 0;
    const int Kx = (int)wshape[2], Ky = (int)wshape[1], Kz = (int)wshape[0];
    if (Kx > 128 || Ky > 128 || Kz > 128) return 0;
    const int lox = Kx / 2 + (int)origins[2];
    const int kxa = Kx + ((lox + 3) / 4 * 4 - lox);
    if (kxa > 16 || (int64_t)Ky * Kz * kxa > (int64_t)(kNdTapBytes / sizeof(float))) return 0;
    int P, TY, TZ;
    if (!nd_static_tile(3, kxa, Ky, Kz, &P, &TY, &TZ)) return 0;
    // the second box of a split tile lands `planes * BY * P` floats into the tile: TMA destinations must
    // be 128-byte aligned, P * 4 is an odd multiple of 16 bytes => the staged rows per plane must be a
    // multiple of 8 (Ky = 1, 9, 17 ... with the 8-row tiles)
    if (((size_t)(TY + Ky - 1) * P * 4) % 128 != 0) return 0;
    const int loz = Kz / 2 + (int)origins[0], hiz = Kz - 1 - loz;
    return nd_peer_tiles_ok(Z, TZ, loz, hiz) ? 1 : 0;
}

template <int KXA, bool SKIPZ>
static int launch_static_cfg(const CorrNdArgs &a, int lox_al)
{
    constexpr int R = 16, SX = 4 * R;
    constexpr int NF = (R + KXA + 3) / 4 * 4;
    const int es = 4;
    const int nd = a.ndim;
    const int64_t X = a.shape[nd - 1], Y = a.shape[nd - 2], Z = nd == 3 ? a.shape[0] : 1;
    const int Kx = (int)a.wshape[nd - 1], Ky = (int)a.wshape[nd - 2], Kz = nd == 3 ? (int)a.wshape[0] : 1;
    const int loy = Ky / 2 + (int)a.origins[nd - 2];
    const int loz = nd == 3 ? Kz / 2 + (int)a.origins[0] : 0;
    const int lead = lox_al - (Kx / 2 + (int)a.origins[nd - 1]);
    if ((int64_t)Ky * Kz * KXA > (int64_t)(kNdTapBytes / sizeof(float))) return -1;

    int P = 0, TY = 0, TZ = 0;
    if (!nd_static_tile(nd, KXA, Ky, Kz, &P, &TY, &TZ)) return -1;
    const int BY = TY + Ky - 1, BZ = TZ + Kz - 1;
    const int64_t cx = (X + SX - 1) / SX, cy = (Y + TY - 1) / TY, cz = (Z + TZ - 1) / TZ;
    const int64_t num_tiles = cx * cy * cz;
    if (num_tiles >= (1ll << 31) || cx >= (1ll << 30) || cy >= (1ll << 30)) return -1;

    const int pad_dim = nd == 3 ? 2 : 1;
    const int64_t Ytot = pad_dim == 1 ? a.pad_lo + Y + a.pad_hi : Y;
    const int64_t Ztot = pad_dim == 2 ? a.pad_lo + Z + a.pad_hi : Z;
    const int64_t plane = pad_dim == 2 ? Y * X : X;
    const char *base = (const char *)a.in - a.pad_lo * plane * es;
    if (X >= (1ll << 31) || Ytot >= (1ll << 31) || Ztot >= (1ll << 31)) return -1;

    CUtensorMap tmap;
    uint64_t dims[3] = {(uint64_t)X, (uint64_t)Ytot, (uint64_t)Ztot};
    uint64_t strides[2] = {(uint64_t)X * es, (uint64_t)X * es * (uint64_t)Ytot};
    uint32_t bx[3] = {(uint32_t)P, (uint32_t)BY, (uint32_t)BZ};
    if (encode_tensor_map(&tmap, a.dtype, (void *)base, 3, dims, strides, bx) != 0) return -1;

    CorrNdTiledParams<float> p;
    p.in = (const float *)a.in; p.out = (float *)a.out;
    p.X = X; p.Y = Y; p.Z = Z;
    p.pad_lo = a.pad_lo; p.pad_hi = a.pad_hi; p.pad_dim = pad_dim;
    p.Kx_al = KXA; p.Ky = Ky; p.Kz = Kz;
    p.lox_al = lox_al; p.loy = loy; p.loz = loz;
    p.TY = TY; p.TZ = TZ; p.ncg = 1; p.P = P; p.BY = BY; p.BZ = BZ;
    p.cx = (int)cx; p.cy = (int)cy;
    p.mode = a.mode; p.cval = (float)a.cval;
    p.wide_store = (((X * es) % 32) == 0 && ((uintptr_t)a.out & 31) == 0) ? 1 : 0;
    const int rows = Ky * Kz;
    for (int i = 0; i < rows * KXA; ++i) p.w[i] = 0.f;
    for (int rw = 0; rw < rows; ++rw)
        for (int kx = 0; kx < Kx; ++kx) {
            const double w = a.w[(size_t)rw * Kx + kx];
            p.w[rw * KXA + lead + kx] = std::fabs(w) > DBL_EPSILON ? (float)w : 0.f;
        }

    const size_t smem = (size_t)P * BY * BZ * es + 128;
    p.peer_lo = p.peer_hi = nullptr; p.n_peer_lo = p.n_peer_hi = 0;
    if (a.peer_lo || a.peer_hi) {
        // neighbour planes read in place: extra tensor maps for the first / last z tile
        const int hiz = Kz - 1 - loz;
        const bool has_lo = a.peer_lo != nullptr && loz > 0, has_hi = a.peer_hi != nullptr && hiz > 0;
        if (nd != 3 || a.pad_lo || a.pad_hi || !nd_peer_tiles_ok(Z, TZ, loz, hiz)) return -1;
        if (((size_t)BY * P * 4) % 128 != 0) return -1;   // 128-byte aligned TMA destinations
        if ((has_lo && a.n_peer_lo != loz) || (has_hi && a.n_peer_hi != hiz)) return -1;
        if (((uintptr_t)a.peer_lo & 15) || ((uintptr_t)a.peer_hi & 15)) return -1;
        NdPeerMaps hm;
        memset(&hm, 0, sizeof(hm));
        const int64_t last_planes = Z - (cz - 1) * TZ;
        hm.has_lo = has_lo; hm.has_hi = has_hi; hm.lo_planes = loz; hm.ncz = (int)cz;
        hm.k_last = (int)(last_planes + loz);
        if (has_lo) {
            uint32_t b1[3] = {(uint32_t)P, (uint32_t)BY, (uint32_t)(BZ - loz)};
            if (encode_tensor_map(&hm.first, a.dtype, (void *)a.in, 3, dims, strides, b1) != 0) return -1;
            uint64_t d2[3] = {(uint64_t)X, (uint64_t)Y, (uint64_t)loz};
            uint32_t b2[3] = {(uint32_t)P, (uint32_t)BY, (uint32_t)loz};
            if (encode_tensor_map(&hm.lo, a.dtype, (void *)a.peer_lo, 3, d2, strides, b2) != 0) return -1;
            p.peer_lo = (const float *)a.peer_lo; p.n_peer_lo = loz; p.pad_lo = loz;
        }
        if (has_hi) {
            uint32_t b1[3] = {(uint32_t)P, (uint32_t)BY, (uint32_t)hm.k_last};
            if (encode_tensor_map(&hm.last, a.dtype, (void *)a.in, 3, dims, strides, b1) != 0) return -1;
            uint64_t d2[3] = {(uint64_t)X, (uint64_t)Y, (uint64_t)hiz};
            uint32_t b2[3] = {(uint32_t)P, (uint32_t)BY, (uint32_t)(BZ - hm.k_last)};
            if (BZ - hm.k_last < 1) return -1;
            if (encode_tensor_map(&hm.hi, a.dtype, (void *)a.peer_hi, 3, d2, strides, b2) != 0) return -1;
            p.peer_hi = (const float *)a.peer_hi; p.n_peer_hi = hiz; p.pad_hi = hiz;
        }
        auto kern = corrnd_static_peer_kernel<KXA, SKIPZ>;
        B200_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemOptIn));
        kern<<<(unsigned)num_tiles, 256, smem, a.stream>>>(tmap, hm, p);
        B200_CUDA_TRY(cudaGetLastError());
        count_launch();
        note_kernel("corrnd_tiled_peer_f32");
        return B200_OK;
    }
    auto kern = corrnd_static_kernel<KXA, SKIPZ>;
    B200_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemOptIn));
    kern<<<(unsigned)num_tiles, 256, smem, a.stream>>>(tmap, p);
    B200_CUDA_TRY(cudaGetLastError());
    count_launch();
    note_kernel("corrnd_tiled_f32");
    return B200_OK;
}

// Does the kernel hold taps the reference skips (|w| <= DBL_EPSILON), or taps that underflow to
// zero in float32?  Such kernels must not multiply them into the sum (non-finite data).
static bool nd_weights_have_zeros(const CorrNdArgs &a)
{
    size_t n = 1;
    for (int d = 0; d < a.ndim; ++d) n *= (size_t)a.wshape[d];
    for (size_t i = 0; i < n; ++i)
        if (!(std::fabs(a.w[i]) > DBL_EPSILON) || (a.dtype == B200_F32 && (float)a.w[i] == 0.f)) return true;
    return false;
}

// float32 with at most 16 taps per weight row (alignment zeros included)
static int launch_static(const CorrNdArgs &a)
{
    const int nd = a.ndim;
    const int64_t X = a.shape[nd - 1];
    if ((X * 4) % 16 != 0) return -1;
    if (((uintptr_t)a.in & 15) || ((uintptr_t)a.out & 15)) return -1;
    const int Kx = (int)a.wshape[nd - 1];
    const int lox = Kx / 2 + (int)a.origins[nd - 1];
    const int lox_al = (lox + 3) / 4 * 4;
    const int kxa = Kx + (lox_al - lox);
    const bool zeros = nd_weights_have_zeros(a);
    switch (kxa) {
#define B200_ND_CASE(K) \
    case K: return zeros ? launch_static_cfg<K, true>(a, lox_al) : launch_static_cfg<K, false>(a, lox_al);
    B200_ND_CASE(1) B200_ND_CASE(2) B200_ND_CASE(3) B200_ND_CASE(4) B200_ND_CASE(5) B200_ND_CASE(6)
    B200_ND_CASE(7) B200_ND_CASE(8) B200_ND_CASE(9) B200_ND_CASE(10) B200_ND_CASE(11) B200_ND_CASE(12)
    B200_ND_CASE(13) B200_ND_CASE(14) B200_ND_CASE(15) B200_ND_CASE(16)
#undef B200_ND_CASE
    default: return -2;   // longer rows: the run-time row engine
    }
}

template <typename T, int SH>
static int launch_sh(const CorrNdArgs &a, int lox_al, int Kx_al)
{
    constexpr int V = 16 / (int)sizeof(T);
    switch (Kx_al % V) {
    case 0: return launch_cfg<T, SH, 0>(a, lox_al, Kx_al);
    case 1: return launch_cfg<T, SH, 1>(a, lox_al, Kx_al);
    case 2: if constexpr (V == 4) return launch_cfg<T, SH, 2>(a, lox_al, Kx_al);
    case 3: if constexpr (V == 4) return launch_cfg<T, SH, 3>(a, lox_al, Kx_al);
    }
    return -1;
}

template <typename T>
static int launch_typed(const CorrNdArgs &a)
{
    constexpr int V = 16 / (int)sizeof(T);
    const int es = (int)sizeof(T);
    const int nd = a.ndim;
    const int64_t X = a.shape[nd - 1];
    if ((X * es) % 16 != 0) return -1;
    if (((uintptr_t)a.in & 15) || ((uintptr_t)a.out & 15)) return -1;
    const int Kx = (int)a.wshape[nd - 1];
    const int64_t rows = a.wshape[nd - 2] * (nd == 3 ? a.wshape[0] : 1);
    const int lox = Kx / 2 + (int)a.origins[nd - 1];
    const int lox_al = (lox + V - 1) / V * V;
    const int shift = lox_al - lox;
    const int Kx_al = Kx + shift;
    if (rows * Kx_al > (int64_t)(kNdTapBytes / sizeof(T))) return -1;
    switch (shift) {
    case 0: return launch_sh<T, 0>(a, lox_al, Kx_al);
    case 1: return launch_sh<T, 1>(a, lox_al, Kx_al);
    case 2: if constexpr (V == 4) return launch_sh<T, 2>(a, lox_al, Kx_al);
    case 3: if constexpr (V == 4) return launch_sh<T, 3>(a, lox_al, Kx_al);
    }
    return -1;
}

// N-D correlate of a slab whose neighbour planes are read in place (a.peer_lo / a.peer_hi); -1 = not eligible
int launch_corrnd_peer(const CorrNdArgs &a)
{
    if (!corrnd_peer_eligible(a.dtype, a.ndim, a.shape, a.wshape, a.origins)) return -1;
    return launch_static(a);
}

int launch_corrnd_tiled(const CorrNdArgs &a)
{
    // eligibility depends on dtype, rank, the contiguous extent and the weights
    // only (never on the number of planes): slabs of one volume agree
    if (a.ndim != 2 && a.ndim != 3) return -1;
    for (int d = 0; d < a.ndim; ++d)
        if (a.wshape[d] > 128) return -1;
    if (a.dtype == B200_F32) {
        // (which engine runs depends on the weights alone)
        const int rc = env_int("B200_ND_STATIC", 1) ? launch_static(a) : -2;
        if (rc != -2) return rc;
    }
    // the run-time row engine multiplies every tap of a row: kernels with zero taps go to the
    // scipy-order kernel (generic.cu), which skips them like the reference
    if (nd_weights_have_zeros(a)) return -1;
    if (a.dtype == B200_F32) return launch_typed<float>(a);
    if (a.dtype == B200_F64) return launch_typed<double>(a);
    return -1;
}

}  // namespace b200
