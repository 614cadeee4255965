// corrnd_int.cuh -- exact tiled N-D correlate for 8- and 16-bit integer arrays
// (2-D and 3-D; convolve / correlate of dask_image.ndfilters on integer images).
//
// scipy's N-D loop (scipy/ndimage/_filters.py:1305 -> NI_Correlate) is
//     tmp = 0;  for every weight with |w| > DBL_EPSILON, in C order:  tmp += w * x
// in double, each product and each sum rounded, then a C cast.  The kernel
// performs those operations in that order with the arithmetic of
// corr1d_int.cuh (2^52-biased FMA product + DADD: two FP64 instructions per
// tap): the array is viewed as [Z][Y][X] (Z = 1 for 2-D), a CTA produces 32
// consecutive y-lines x `tile_out` x-positions of one z-plane out of a shared
// tile of KZ x (32 + KY - 1) staged lines, lane l owns line l and walks along x
// four outputs at a time, one weight row (kz, ky) after the other.  Weights the
// reference skips are carried as zeros (x * 0 adds +0: same sum).  The phases
// are __host__ __device__ (tests/native/int_emul.cu runs them on the CPU).
#pragma once
#include "corr1d_int.cuh"
#include <vector>

namespace b200 {

constexpr int kIntNdLines = 32;        // output lines per tile, one per lane
constexpr int kIntNdMaxTaps = 1024;
constexpr size_t kIntNdMaxSmem = 96 * 1024;

struct IntNdParams {
    const void *in;
    void *out;
    const IntTap *taps;     // KZ * KY * KX {w, c} pairs, C order (device memory)
    int64_t Z, Y, X;
    int64_t pad_lo, pad_hi;
    int pad_axis;           // 0: the pads extend Z, 1: they extend Y (2-D arrays)
    int KZ, KY, KX, ntaps;
    int loz, loy, lox;      // tap k of an axis reads position i - lo + k
    int mode, cval;
    int tile_out;           // x-positions per tile (multiple of 16)
    int tile_len;           // staged positions per line
    int shift, pitch, opitch;
    int stage_lines;        // 32 + KY - 1
    int vec;
    int c_tiles, y_tiles;
    uint32_t hi_word;
    uint32_t vpr_magic, ly_magic;   // ceil(2^32 / d) for d = vectors per line, staged lines per plane
};

struct IntNdTile {
    int64_t z, y0, x0, g0;
    __host__ __device__ __forceinline__ IntNdTile(const IntNdParams &p, uint32_t bid)
    {
        // (the grid is below 2^31 blocks: 32-bit divisions)
        const uint32_t q = bid / (uint32_t)p.c_tiles;
        const int tcx = (int)(bid - q * (uint32_t)p.c_tiles);
        const uint32_t zz = q / (uint32_t)p.y_tiles;
        y0 = (int64_t)(q - zz * (uint32_t)p.y_tiles) * kIntNdLines;
        z = zz;
        x0 = (int64_t)tcx * p.tile_out;
        g0 = x0 - p.lox - p.shift;
    }
};

__host__ __device__ __forceinline__ size_t int_nd_tap_bytes(int ntaps) { return ((size_t)ntaps * 16 + 127) / 128 * 128; }
__host__ __device__ __forceinline__ size_t int_nd_table_bytes(const IntNdParams &p)
{
    return ((size_t)p.KZ * p.stage_lines * 8 + 127) / 128 * 128;
}

// Source of every staged line: element offset of its row (z, y) in the array, or
// INT64_MIN when the row lies outside under mode 'constant'.  One entry per line, computed
// once per tile (the extension map costs 64-bit compares and, on the edges, a modulo).
__host__ __device__ __forceinline__ void int_nd_stage_lines(const IntNdParams &p, int64_t *table, int tid, uint32_t bid)
{
    const IntNdTile t(p, bid);
    const int LY = p.stage_lines;
    for (int line = tid; line < p.KZ * LY; line += kIntThreads) {
        const int lz = (int)int_div_magic((uint32_t)line, p.ly_magic), ly = line - lz * LY;
        const int64_t zpos = t.z - p.loz + lz, ypos = t.y0 - p.loy + ly;
        const int64_t zs = p.pad_axis == 0 ? ext_index_padded(zpos, p.Z, p.mode, p.pad_lo, p.pad_hi)
                                           : ext_index_padded(zpos, p.Z, p.mode, 0, 0);
        const int64_t ys = p.pad_axis == 1 ? ext_index_padded(ypos, p.Y, p.mode, p.pad_lo, p.pad_hi)
                                           : ext_index_padded(ypos, p.Y, p.mode, 0, 0);
        // (a padded slab addresses rows in front of `in`: offsets may be negative, the flag is separate)
        table[line] = (zs == B200_USE_CVAL || ys == B200_USE_CVAL) ? INT64_MIN : (zs * p.Y + ys) * p.X;
    }
}

__host__ __device__ __forceinline__ void int_nd_stage_taps(const IntNdParams &p, IntTap *wc, int tid)
{
    for (int i = tid; i < p.ntaps; i += kIntThreads) wc[i] = p.taps[i];
}

template <typename TS>
__host__ __device__ __forceinline__ void int_nd_stage(const IntNdParams &p, TS *tile, const int64_t *table, int tid,
                                                      uint32_t bid)
{
    constexpr int A = 16 / (int)sizeof(TS);
    const IntNdTile t(p, bid);
    const TS *in = (const TS *)p.in;
    const int PC = p.tile_len, P = p.pitch;
    const int VPR = PC / A;
    const int nvec = p.KZ * p.stage_lines * VPR;
    TS fill[A];
#pragma unroll
    for (int e = 0; e < A; ++e) fill[e] = (TS)p.cval;
    uint4 cvec;
    memcpy(&cvec, fill, 16);
    for (int v = tid; v < nvec; v += kIntThreads) {
        const int line = (int)int_div_magic((uint32_t)v, p.vpr_magic), cv = v - line * VPR;
        const int64_t off = table[line];
        const int64_t pos = t.g0 + cv * A;
        uint4 val = cvec;
        if (off != INT64_MIN) {
            const TS *src = in + off;
            if (p.vec && pos >= 0 && pos + A <= p.X) {
                val = *reinterpret_cast<const uint4 *>(src + pos);
            } else {
                TS el[A];
#pragma unroll
                for (int e = 0; e < A; ++e) {
                    const int64_t xs = ext_index(pos + e, p.X, p.mode);
                    el[e] = xs < 0 ? (TS)p.cval : src[xs];
                }
                memcpy(&val, el, 16);
            }
        }
        // lines start on 4-byte, not 16-byte, boundaries
        uint32_t *d = reinterpret_cast<uint32_t *>(tile + line * P + cv * A);
        d[0] = val.x; d[1] = val.y; d[2] = val.z; d[3] = val.w;
    }
}

// KXS = taps per weight row when it is one of the short static cases, 0 = run-time loop
template <typename TS, int KXS>
__host__ __device__ __forceinline__ void int_nd_compute_kx(const IntNdParams &p, const TS *tile, TS *otile,
                                                           const IntTap *wc, int tid)
{
    const int warp = tid >> 5, lane = tid & 31;
    const int per_warp = p.tile_out / (kIntThreads / 32);   // multiple of 4
    const int P = p.pitch, LY = p.stage_lines;
    TS *oline = otile + lane * p.opitch;
    for (int g = 0; g < per_warp; g += 4) {
        const int xo = warp * per_warp + g;
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
        const IntTap *wrow = wc;
        for (int kz = 0; kz < p.KZ; ++kz) {
            const TS *row = tile + (kz * LY + lane) * P + p.shift + xo;
            for (int ky = 0; ky < p.KY; ++ky, row += P, wrow += p.KX) {
                if constexpr (KXS > 0) int_quad_taps_static<TS, KXS>(row, wrow, p.hi_word, acc);
                else int_quad_taps<TS>(row, 1, p.KX, wrow, p.hi_word, acc);
            }
        }
        TS r[4];
#pragma unroll
        for (int m = 0; m < 4; ++m) r[m] = cast_out<TS>(acc[m]);
        uint32_t *d = reinterpret_cast<uint32_t *>(oline + xo);
        if constexpr (sizeof(TS) == 2) {
            d[0] = (uint32_t)(uint16_t)r[0] | ((uint32_t)(uint16_t)r[1] << 16);
            d[1] = (uint32_t)(uint16_t)r[2] | ((uint32_t)(uint16_t)r[3] << 16);
        } else {
            d[0] = (uint32_t)(uint8_t)r[0] | ((uint32_t)(uint8_t)r[1] << 8) | ((uint32_t)(uint8_t)r[2] << 16) |
                   ((uint32_t)(uint8_t)r[3] << 24);
        }
    }
}

template <typename TS>
__host__ __device__ __forceinline__ void int_nd_compute(const IntNdParams &p, const TS *tile, TS *otile,
                                                        const IntTap *wc, int tid, uint32_t)
{
    switch (p.KX) {   // uniform over the grid
    case 1: int_nd_compute_kx<TS, 1>(p, tile, otile, wc, tid); break;
    case 2: int_nd_compute_kx<TS, 2>(p, tile, otile, wc, tid); break;
    case 3: int_nd_compute_kx<TS, 3>(p, tile, otile, wc, tid); break;
    case 4: int_nd_compute_kx<TS, 4>(p, tile, otile, wc, tid); break;
    case 5: int_nd_compute_kx<TS, 5>(p, tile, otile, wc, tid); break;
    case 7: int_nd_compute_kx<TS, 7>(p, tile, otile, wc, tid); break;
    default: int_nd_compute_kx<TS, 0>(p, tile, otile, wc, tid); break;
    }
}

template <typename TS>
__host__ __device__ __forceinline__ void int_nd_writeout(const IntNdParams &p, const TS *otile, int tid, uint32_t bid)
{
    constexpr int A = 16 / (int)sizeof(TS);
    const IntNdTile t(p, bid);
    TS *plane = (TS *)p.out + t.z * p.Y * p.X;
    const int SX = p.tile_out, PO = p.opitch;
    if (p.vec) {
        const int VR = SX / A;   // a power of two
        const int vr_shift = 31 - __builtin_clz((unsigned)VR);
        for (int v = tid; v < kIntNdLines * VR; v += kIntThreads) {
            const int r = v >> vr_shift, cv = v - r * VR;
            const int64_t y = t.y0 + r, x = t.x0 + cv * A;
            if (y >= p.Y || x >= p.X) continue;
            const uint32_t *sw = reinterpret_cast<const uint32_t *>(otile + r * PO + cv * A);
            *reinterpret_cast<uint4 *>(plane + y * p.X + x) = make_uint4(sw[0], sw[1], sw[2], sw[3]);
        }
    } else {
        for (int e = tid; e < kIntNdLines * SX; e += kIntThreads) {
            const int r = e / SX, cc = e - r * SX;
            const int64_t y = t.y0 + r, x = t.x0 + cc;
            if (y >= p.Y || x >= p.X) continue;
            plane[y * p.X + x] = otile[r * PO + cc];
        }
    }
}

// ---------------------------------------------------------------------------
// host: eligibility, geometry and the tap table.  Returns 0 when eligible.
// ---------------------------------------------------------------------------
struct IntNdPlan {
    IntNdParams p;
    std::vector<IntTap> taps;
    size_t smem_bytes;
    int64_t blocks;
};

inline int plan_corrnd_int(const CorrNdArgs &a, IntNdPlan &pl)
{
    const int dt = a.dtype;
    if (dt != B200_U8 && dt != B200_I8 && dt != B200_U16 && dt != B200_I16) return -1;
    if (a.ndim != 2 && a.ndim != 3) return -1;
    const int sz = dtype_size(dt);
    const int A = 16 / sz;
    IntNdParams &p = pl.p;
    const int o = a.ndim == 3 ? 0 : -1;   // array axis of Z
    p.Z = o >= 0 ? a.shape[0] : 1;
    p.Y = a.shape[o + 1];
    p.X = a.shape[o + 2];
    p.KZ = o >= 0 ? (int)a.wshape[0] : 1;
    p.KY = (int)a.wshape[o + 1];
    p.KX = (int)a.wshape[o + 2];
    const int64_t ntaps = (int64_t)p.KZ * p.KY * p.KX;
    if (ntaps > kIntNdMaxTaps || p.KY > 64) return -1;
    p.ntaps = (int)ntaps;
    p.loz = o >= 0 ? (int)(a.wshape[0] / 2 + a.origins[0]) : 0;
    p.loy = (int)(a.wshape[o + 1] / 2 + a.origins[o + 1]);
    p.lox = (int)(a.wshape[o + 2] / 2 + a.origins[o + 2]);
    p.pad_lo = a.pad_lo; p.pad_hi = a.pad_hi;
    p.pad_axis = a.ndim == 3 ? 0 : 1;
    p.mode = a.mode;
    p.cval = 0;
    p.hi_word = kIntHigh;
    if (a.mode == B200_MODE_CONSTANT) {
        const double lo_v = dt == B200_I8 ? -128.0 : (dt == B200_I16 ? -32768.0 : 0.0);
        const double hi_v = dt == B200_U8 ? 255.0 : (dt == B200_I8 ? 127.0 : (dt == B200_U16 ? 65535.0 : 32767.0));
        if (!(a.cval >= lo_v && a.cval <= hi_v) || a.cval != std::floor(a.cval)) return -1;
        p.cval = (int)a.cval;
    }
    if (p.X >= (1ll << 31) - 4096 || p.Y >= (1ll << 31) - 4096) return -1;
    pl.taps.resize(p.ntaps);
    for (int i = 0; i < p.ntaps; ++i) {
        double w = a.w[i];
        if (!std::isfinite(w) || std::fabs(w) >= 0x1p900) return -1;
        if (!(std::fabs(w) > DBL_EPSILON)) w = 0.0;   // the reference leaves these taps out
        pl.taps[i] = IntTap{w, -(0x1p52 * w)};
    }
    p.in = a.in; p.out = a.out;
    p.taps = nullptr;
    p.vec = (p.X % A == 0) && ((uintptr_t)a.in % 16 == 0) && ((uintptr_t)a.out % 16 == 0);
    p.stage_lines = kIntNdLines + p.KY - 1;
    p.shift = (A - p.lox % A) % A;
    for (p.tile_out = 128; p.tile_out >= 32; p.tile_out /= 2) {
        p.tile_len = (p.shift + p.tile_out + p.KX + A - 1) / A * A;
        p.pitch = p.tile_len + 4 / sz;
        p.opitch = p.tile_out + 4 / sz;
        pl.smem_bytes = int_nd_tap_bytes(p.ntaps) + int_nd_table_bytes(p) +
                        ((size_t)p.KZ * p.stage_lines * p.pitch + (size_t)kIntNdLines * p.opitch) * sz;
        if (pl.smem_bytes <= kIntNdMaxSmem) break;
    }
    if (p.tile_out < 32) return -1;
    const int64_t nvec = (int64_t)p.KZ * p.stage_lines * (p.tile_len / A);
    if (nvec >= 65536) return -1;
    p.vpr_magic = (uint32_t)(((1ull << 32) + (p.tile_len / A) - 1) / (p.tile_len / A));
    p.ly_magic = (uint32_t)(((1ull << 32) + p.stage_lines - 1) / p.stage_lines);
    p.c_tiles = (int)((p.X + p.tile_out - 1) / p.tile_out);
    p.y_tiles = (int)((p.Y + kIntNdLines - 1) / kIntNdLines);
    pl.blocks = p.Z * (int64_t)p.y_tiles * p.c_tiles;
    if (pl.blocks < 1 || pl.blocks >= (1ll << 31)) return -1;
    return 0;
}

}  // namespace b200
