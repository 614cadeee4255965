// corr1d_int.cuh -- exact tiled correlate1d for 8- and 16-bit integer arrays.
//
// scipy runs every 1-D pass of an integer array in double precision and casts
// the sum back with C truncation (scipy/ndimage/_filters.py:610 ->
// NI_Correlate1D), so an integer gaussian is only reproducible bit for bit by
// performing the same double operations in the same order:
//
//   symmetric taps      acc = x[c]*w[c];  acc += (x[c-j] + x[c+j]) * w[c-j]   j = size1 .. 1
//   antisymmetric taps  acc = x[c]*w[c];  acc += (x[c-j] - x[c+j]) * w[c-j]
//   otherwise           acc = x[last]*w[last];  acc += x[k]*w[k]              k = 0 .. L-2
//
// each product rounded, then each sum rounded.  The kernels here do exactly
// that, but out of a shared-memory tile instead of one global load per tap, and
// with two FP64 instructions per tap (pair) instead of four:
//
//   * x[c-j] +- x[c+j] is formed in INTEGER arithmetic (exact);
//   * the integer v becomes the double 2^52 + |v| by pairing it with the
//     constant high word 0x43300000 -- no I2F;
//   * fma(2^52 + |v|, w, -(2^52 * w)) has the exact value |v| * w before its
//     single rounding (2^52 * w is a power-of-two scaling, hence exact), so it
//     IS the rounded product |v| * w; the sign of v is XORed into the result
//     (round-to-nearest is symmetric);
//   * a separate DADD accumulates, as the reference does.
//
// Every lane owns one LINE of the tile and walks along it four consecutive
// outputs at a time, keeping the left and right ends of the tap pairs in two
// rotating 4-element register windows: one shared-memory load per window per
// tap pair serves four outputs.  The functions are __host__ __device__ and
// split in the phases a CTA separates with __syncthreads(), so that the CPU test
// suite can run the very same code thread by thread against the oracle
// (tests/native/int_emul.cu).
#pragma once
#include "common.cuh"
#include <cstring>
#include <cmath>
#include <type_traits>

namespace b200 {

constexpr int kMaxTapsInt = 129;
constexpr int kIntThreads = 128;
constexpr int kIntStridedW = 128;   // strided tiles: columns per tile, one per thread
constexpr int kIntContigRows = 32;  // contiguous tiles: lines per tile, one per lane
constexpr int kIntContigOut = 128;  // contiguous tiles: outputs per line
constexpr int kIntTapBytes = ((kMaxTapsInt * 16 + 127) / 128) * 128;   // shared {w, c} table in front of the tile

struct IntParams {
    const void *in;
    void *out;
    int64_t outer, n, inner;
    int64_t pad_lo, pad_hi;
    int nw, lo, mode, epilogue;
    int nsteps;     // tap pairs (symmetric / antisymmetric) or L - 1 single taps
    int centre;     // index of the tap that opens the sum
    int tile_out;   // outputs per tile along the filtered axis (multiple of 4; contiguous: of 16)
    int tile_len;   // staged positions per line
    int shift;      // contiguous: positions staged in front of the first tap (16-byte aligned loads)
    int pitch;      // contiguous: elements between lines of the staged tile
    int opitch;     // contiguous: elements between lines of the output tile
    int vec;        // 16-byte global accesses are legal
    int n_tiles, c_tiles;
    int cval;
    uint32_t vpr_magic;   // contiguous: ceil(2^32 / vectors per staged line)
    uint32_t hi_word;   // kIntHigh (see int_pattern)
    double w[kMaxTapsInt];
    double c[kMaxTapsInt];   // -(2^52 * w)
};

// ---------------------------------------------------------------------------
// the arithmetic
// ---------------------------------------------------------------------------
__host__ __device__ __forceinline__ double int_make_double(uint32_t hi, uint32_t lo)
{
#ifdef __CUDA_ARCH__
    return __hiloint2double((int)hi, (int)lo);
#else
    const uint64_t b = ((uint64_t)hi << 32) | lo;
    double d;
    memcpy(&d, &b, 8);
    return d;
#endif
}

__host__ __device__ __forceinline__ double int_xor_sign(double t, uint32_t signbit)
{
#ifdef __CUDA_ARCH__
    return __hiloint2double(__double2hiint(t) ^ (int)signbit, __double2loint(t));
#else
    uint64_t b;
    memcpy(&b, &t, 8);
    b ^= (uint64_t)signbit << 32;
    memcpy(&t, &b, 8);
    return t;
#endif
}

__host__ __device__ __forceinline__ double int_fma(double a, double b, double c)
{
#ifdef __CUDA_ARCH__
    return __fma_rn(a, b, c);
#else
    return std::fma(a, b, c);
#endif
}

__host__ __device__ __forceinline__ double int_add(double a, double b)
{
#ifdef __CUDA_ARCH__
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}

// v / d for v < 2^16 with magic = ceil(2^32 / d), d <= 2^16 (exact: the error term v * (magic * d - 2^32)
// stays below 2^32)
__host__ __device__ __forceinline__ uint32_t int_div_magic(uint32_t v, uint32_t magic)
{
#ifdef __CUDA_ARCH__
    return __umulhi(v, magic);
#else
    return (uint32_t)(((uint64_t)v * magic) >> 32);
#endif
}

constexpr uint32_t kIntHigh = 0x43300000u;   // high word of the doubles in [2^52, 2^53)

// round(v * w) for an integer |v| < 2^32: c = -(2^52 * w)
template <bool SIGNEDV>
__host__ __device__ __forceinline__ double int_product(int v, double w, double c)
{
    if constexpr (SIGNEDV) {
        const uint32_t a = (uint32_t)(v < 0 ? -v : v);
        const double t = int_fma(int_make_double(kIntHigh, a), w, c);
        return int_xor_sign(t, (uint32_t)v & 0x80000000u);
    } else {
        return int_fma(int_make_double(kIntHigh, (uint32_t)v), w, c);
    }
}

// Unsigned single taps (no pair to add): a window element is kept as the BIT
// PATTERN of 2^52 + x (the element in the low word, kIntHigh above it) and
// feeds the DFMA of four consecutive steps as it is.
// (`hi` is kIntHigh carried in the kernel parameters: as a literal, ptxas folds
// it back out of the register pair and re-creates it in front of every use.)
__host__ __device__ __forceinline__ uint64_t int_pattern(uint32_t hi, uint32_t x)
{
#ifdef __CUDA_ARCH__
    uint64_t q;
    asm volatile("mov.b64 %0, {%1, %2};" : "=l"(q) : "r"(x), "r"(hi));
    return q;
#else
    return ((uint64_t)hi << 32) | x;
#endif
}

__host__ __device__ __forceinline__ double int_pattern_product(uint64_t sq, double w, double c)
{
#ifdef __CUDA_ARCH__
    return __fma_rn(__longlong_as_double((long long)sq), w, c);
#else
    double d;
    memcpy(&d, &sq, 8);
    return std::fma(d, w, c);
#endif
}

// One tap: the weight and c = -(2^52 * w).  The tap loop reads the pair from a
// shared-memory table with one broadcast LDS.128 (an indexed read of the kernel
// parameters is two LDC.64).
struct alignas(16) IntTap {
    double w, c;
};

__host__ __device__ __forceinline__ void int_stage_taps(const IntParams &p, IntTap *wc, int tid)
{
    for (int i = tid; i < p.nw; i += kIntThreads) wc[i] = IntTap{p.w[i], p.c[i]};
}

// Steps 0 .. ns-1 through `one.operator()<U>(k)` with U = k % NQ static: the
// window slots of the callers rotate with period NQ (the outputs per lane).
template <int NQ, typename F>
__host__ __device__ __forceinline__ void int_run_steps(int ns, F &&one)
{
    static_assert(NQ == 4 || NQ == 8, "four or eight outputs per lane");
    int k = 0;
    for (; k + NQ <= ns; k += NQ) {
        one.template operator()<0>(k);
        one.template operator()<1>(k + 1);
        one.template operator()<2>(k + 2);
        one.template operator()<3>(k + 3);
        if constexpr (NQ == 8) {
            one.template operator()<4>(k + 4);
            one.template operator()<5>(k + 5);
            one.template operator()<6>(k + 6);
            one.template operator()<7>(k + 7);
        }
    }
    if (k < ns) one.template operator()<0>(k);
    if (k + 1 < ns) one.template operator()<1>(k + 1);
    if (k + 2 < ns) one.template operator()<2>(k + 2);
    if constexpr (NQ == 8) {
        if (k + 3 < ns) one.template operator()<3>(k + 3);
        if (k + 4 < ns) one.template operator()<4>(k + 4);
        if (k + 5 < ns) one.template operator()<5>(k + 5);
        if (k + 6 < ns) one.template operator()<6>(k + 6);
    }
}

// acc[m] += x[m + k] * w[k] for k = 0 .. ns-1, in that order, for NQ
// consecutive outputs m: the single-tap walk (general 1-D weights, and every
// weight row of the N-D correlate).  `s` addresses x[0], `step` is the element
// distance between consecutive positions.  Reads s[0 .. (ns+NQ-1)*step].
template <typename TS, int NQ = 4>
__host__ __device__ __forceinline__ void int_quad_taps(const TS *s, int step, int ns, const IntTap *wc, uint32_t hi,
                                                       double (&acc)[NQ])
{
    const TS *sa = s + NQ * step;   // next element entering the window
    if constexpr (std::is_signed<TS>::value) {
        int A[NQ];
#pragma unroll
        for (int m = 0; m < NQ; ++m) A[m] = (int)s[m * step];
        int_run_steps<NQ>(ns, [&]<int U>(int k) {
            const IntTap q = wc[k];
#pragma unroll
            for (int m = 0; m < NQ; ++m)
                acc[m] = int_add(acc[m], int_product<true>(A[(U + m) & (NQ - 1)], q.w, q.c));
            A[U] = (int)*sa;
            sa += step;
        });
    } else {
        uint64_t A[NQ];
#pragma unroll
        for (int m = 0; m < NQ; ++m) A[m] = int_pattern(hi, (uint32_t)s[m * step]);
        int_run_steps<NQ>(ns, [&]<int U>(int k) {
            const IntTap q = wc[k];
#pragma unroll
            for (int m = 0; m < NQ; ++m)
                acc[m] = int_add(acc[m], int_pattern_product(A[(U + m) & (NQ - 1)], q.w, q.c));
            A[U] = int_pattern(hi, (uint32_t)*sa);
            sa += step;
        });
    }
}

// The same walk with the tap count known at compile time (short weight rows of
// the N-D correlate): the NT + 3 elements are loaded once, no loop, no window
// rotation.
template <typename TS, int NT>
__host__ __device__ __forceinline__ void int_quad_taps_static(const TS *s, const IntTap *wc, uint32_t hi,
                                                              double (&acc)[4])
{
    if constexpr (std::is_signed<TS>::value) {
        int x[NT + 3];
#pragma unroll
        for (int i = 0; i < NT + 3; ++i) x[i] = (int)s[i];
#pragma unroll
        for (int k = 0; k < NT; ++k) {
            const IntTap q = wc[k];
#pragma unroll
            for (int m = 0; m < 4; ++m) acc[m] = int_add(acc[m], int_product<true>(x[m + k], q.w, q.c));
        }
    } else {
        uint64_t x[NT + 3];
#pragma unroll
        for (int i = 0; i < NT + 3; ++i) x[i] = int_pattern(hi, (uint32_t)s[i]);
#pragma unroll
        for (int k = 0; k < NT; ++k) {
            const IntTap q = wc[k];
#pragma unroll
            for (int m = 0; m < 4; ++m) acc[m] = int_add(acc[m], int_pattern_product(x[m + k], q.w, q.c));
        }
    }
}

// KIND: +1 symmetric, -1 antisymmetric, 0 general.  `s` addresses the element
// under tap 0 of the first of NQ consecutive outputs, `step` is the element
// distance between consecutive positions of the line.  Reads s[0 .. (nw+NQ-2)*step].
// Element q of the line lives in window slot A[q % NQ]; element (L - 1 + q') in
// B[q' % NQ]; the tap loop is unrolled NQ times so that the slots are static.
template <typename TS, int KIND, int NQ = 4>
__host__ __device__ __forceinline__ void int_line_quad(const TS *s, int step, const IntParams &p, const IntTap *wc,
                                                       double (&acc)[NQ])
{
    constexpr bool SV = KIND < 0 || std::is_signed<TS>::value;
    const int L = p.nw;
    {
        const IntTap q = wc[p.centre];
        const TS *sc = s + p.centre * step;
#pragma unroll
        for (int m = 0; m < NQ; ++m) acc[m] = int_product<std::is_signed<TS>::value>((int)sc[m * step], q.w, q.c);
    }
    if constexpr (KIND == 0) {
        int_quad_taps<TS, NQ>(s, step, p.nsteps, wc, p.hi_word, acc);
    } else {
        const TS *sa = s + NQ * step;        // next element entering the left window
        const TS *sb = s + (L - 2) * step;   // next element entering the right window
        int A[NQ], B[NQ];
#pragma unroll
        for (int m = 0; m < NQ; ++m) {
            A[m] = (int)s[m * step];
            B[m] = (int)s[(L - 1 + m) * step];
        }
        int_run_steps<NQ>(p.nsteps, [&]<int U>(int k) {
            const IntTap q = wc[k];
#pragma unroll
            for (int m = 0; m < NQ; ++m) {
                const int a = A[(U + m) & (NQ - 1)];
                const int b = B[(m - U) & (NQ - 1)];
                acc[m] = int_add(acc[m], int_product<SV>(KIND > 0 ? a + b : a - b, q.w, q.c));
            }
            A[U] = (int)*sa;
            sa += step;
            B[(NQ - 1 - U) & (NQ - 1)] = (int)*sb;
            sb -= step;
        });
    }
}

template <typename TS>
__host__ __device__ __forceinline__ TS int_finish(double acc, TS *dst, int epilogue)
{
    TS r = cast_out<TS>(acc);
    if (epilogue != B200_EPI_STORE) {
        const TS prev = epilogue_reads_out(epilogue) ? *dst : TS(0);
        r = apply_epilogue<TS>(epilogue, prev, r);
    }
    return r;
}

// ---------------------------------------------------------------------------
// strided axis: the array is [outer][n][inner], the filter runs along n.  A
// tile is tile_len = tile_out + nw - 1 rows of kIntStridedW contiguous columns;
// thread `tid` owns column tid.
// ---------------------------------------------------------------------------
struct IntStridedTile {
    int64_t o, row0, col0, g0;
    __host__ __device__ __forceinline__ IntStridedTile(const IntParams &p, uint32_t bid)
    {
        // (the grid is below 2^31 blocks: 32-bit divisions)
        const uint32_t q = bid / (uint32_t)p.n_tiles;
        const int tn = (int)(bid - q * (uint32_t)p.n_tiles);
        const uint32_t oo = q / (uint32_t)p.c_tiles;
        const int tc = (int)(q - oo * (uint32_t)p.c_tiles);
        o = oo;
        row0 = (int64_t)tn * p.tile_out;
        col0 = (int64_t)tc * kIntStridedW;
        g0 = row0 - p.lo;
    }
};

// Source of every staged row: element offset of the row in the array (negative
// for the planes a padded slab keeps in front of `in`), or INT64_MIN for "cval".
// Computed once per tile: the extension map is 64-bit compares and, on the
// array's edges, a modulo.
constexpr int kIntRowTableBytes = (32 + kMaxTapsInt + 15) / 16 * 16 * 8;

__host__ __device__ __forceinline__ void int_strided_stage_rows(const IntParams &p, int64_t *rows, int tid, uint32_t bid)
{
    const IntStridedTile t(p, bid);
    for (int r = tid; r < p.tile_len; r += kIntThreads) {
        const int64_t src = ext_index_padded(t.g0 + r, p.n, p.mode, p.pad_lo, p.pad_hi);
        rows[r] = src == B200_USE_CVAL ? INT64_MIN : (t.o * p.n + src) * p.inner;
    }
}

template <typename TS>
__host__ __device__ __forceinline__ void int_strided_stage(const IntParams &p, TS *tile, const int64_t *rows, int tid,
                                                           uint32_t bid)
{
    constexpr int W = kIntStridedW;
    constexpr int A = 16 / (int)sizeof(TS);
    const IntStridedTile t(p, bid);
    const TS *in = (const TS *)p.in;
    const int SR = p.tile_len;
    if (p.vec) {
        constexpr int VPR = W / A;
        TS fill[A];
#pragma unroll
        for (int e = 0; e < A; ++e) fill[e] = (TS)p.cval;
        uint4 cvec;
        memcpy(&cvec, fill, 16);
        for (int v = tid; v < SR * VPR; v += kIntThreads) {
            const int r = v / VPR, cv = v - r * VPR;
            const int64_t col = t.col0 + cv * A;
            uint4 val = make_uint4(0u, 0u, 0u, 0u);
            if (col < p.inner) {
                const int64_t off = rows[r];
                val = off == INT64_MIN ? cvec : *reinterpret_cast<const uint4 *>(in + off + col);
            }
            *reinterpret_cast<uint4 *>(tile + r * W + cv * A) = val;
        }
    } else {
        for (int e = tid; e < SR * W; e += kIntThreads) {
            const int r = e / W, cc = e - r * W;
            const int64_t col = t.col0 + cc;
            TS val = TS(0);
            if (col < p.inner) {
                const int64_t off = rows[r];
                val = off == INT64_MIN ? (TS)p.cval : in[off + col];
            }
            tile[e] = val;
        }
    }
}

// NQ = outputs per lane and walk (4, or 8 in the B200_INT_NQ8 build)
template <typename TS, int KIND, int NQ = 4>
__host__ __device__ __forceinline__ void int_strided_compute(const IntParams &p, const TS *tile, const IntTap *wc,
                                                             int tid, uint32_t bid)
{
    constexpr int W = kIntStridedW;
    const IntStridedTile t(p, bid);
    // (no early exit for the columns past the array: the tap loop stays convergent,
    // which keeps the weights on the uniform datapath; only the stores are predicated)
    const int64_t col = t.col0 + tid;
    const bool live = col < p.inner;
    const bool plain = p.epilogue == B200_EPI_STORE;
    TS *dst = (TS *)p.out + (t.o * p.n + t.row0) * p.inner + col;
    const int64_t left = p.n - t.row0;
    const int rows_here = left < p.tile_out ? (int)left : p.tile_out;   // rows of the array in this tile
    for (int r0 = 0; r0 < rows_here; r0 += NQ) {
        double acc[NQ];
        int_line_quad<TS, KIND, NQ>(tile + r0 * W + tid, W, p, wc, acc);
#pragma unroll
        for (int m = 0; m < NQ; ++m, dst += p.inner) {
            if (live && r0 + m < rows_here)
                *dst = plain ? cast_out<TS>(acc[m]) : int_finish<TS>(acc[m], dst, p.epilogue);
        }
    }
}

// ---------------------------------------------------------------------------
// contiguous axis: the array is [outer][n], the filter runs along n.  A tile is
// kIntContigRows lines of tile_len staged positions, `pitch` elements apart with
// pitch * sizeof(TS) / 4 odd, so that the 32 lanes of a warp (one line each)
// read 32 different banks.  Tile column 0 holds position x0 - lo - shift.
// Results go through a second shared tile so that global stores are row-wise.
// ---------------------------------------------------------------------------
struct IntContigTile {
    int64_t row0, x0, g0;
    __host__ __device__ __forceinline__ IntContigTile(const IntParams &p, uint32_t bid)
    {
        const uint32_t q = bid / (uint32_t)p.c_tiles;
        const int tcx = (int)(bid - q * (uint32_t)p.c_tiles);
        row0 = (int64_t)q * kIntContigRows;
        x0 = (int64_t)tcx * p.tile_out;
        g0 = x0 - p.lo - p.shift;
    }
};

template <typename TS>
__host__ __device__ __forceinline__ void int_contig_stage(const IntParams &p, TS *tile, int tid, uint32_t bid)
{
    constexpr int TR = kIntContigRows;
    constexpr int A = 16 / (int)sizeof(TS);
    const IntContigTile t(p, bid);
    const TS *in = (const TS *)p.in;
    const int PC = p.tile_len, P = p.pitch;
    const int VPR = PC / A;
    for (int v = tid; v < TR * VPR; v += kIntThreads) {
        const int r = (int)int_div_magic((uint32_t)v, p.vpr_magic), cv = v - r * VPR;
        const int64_t row = t.row0 + r;
        const int64_t pos = t.g0 + cv * A;
        uint4 val = make_uint4(0u, 0u, 0u, 0u);
        if (row < p.outer) {
            const TS *line = in + row * p.n;
            if (p.vec && pos >= 0 && pos + A <= p.n) {
                val = *reinterpret_cast<const uint4 *>(line + pos);
            } else {
                TS el[A];
#pragma unroll
                for (int e = 0; e < A; ++e) {
                    const int64_t src = ext_index(pos + e, p.n, p.mode);
                    el[e] = src < 0 ? (TS)p.cval : line[src];
                }
                memcpy(&val, el, 16);
            }
        }
        // lines start on 4-byte, not 16-byte, boundaries
        uint32_t *d = reinterpret_cast<uint32_t *>(tile + r * P + cv * A);
        d[0] = val.x; d[1] = val.y; d[2] = val.z; d[3] = val.w;
    }
}

template <typename TS, int KIND, int NQ = 4>
__host__ __device__ __forceinline__ void int_contig_compute(const IntParams &p, const TS *tile, TS *otile,
                                                            const IntTap *wc, int tid, uint32_t bid)
{
    const IntContigTile t(p, bid);
    const int warp = tid >> 5, lane = tid & 31;
    const int per_warp = p.tile_out / (kIntThreads / 32);   // multiple of 8
    const TS *line = tile + lane * p.pitch + p.shift;
    TS *oline = otile + lane * p.opitch;
    for (int g = 0; g < per_warp; g += NQ) {
        // (groups past the end of a ragged line are computed and dropped: a per-warp
        // exit would take the weights off the uniform datapath)
        const int xo = warp * per_warp + g;
        double acc[NQ];
        int_line_quad<TS, KIND, NQ>(line + xo, 1, p, wc, acc);
        TS r[NQ];
#pragma unroll
        for (int m = 0; m < NQ; ++m) r[m] = cast_out<TS>(acc[m]);
        uint32_t *d = reinterpret_cast<uint32_t *>(oline + xo);
#pragma unroll
        for (int h = 0; h < NQ / 4; ++h) {
            if constexpr (sizeof(TS) == 2) {
                d[2 * h] = (uint32_t)(uint16_t)r[4 * h] | ((uint32_t)(uint16_t)r[4 * h + 1] << 16);
                d[2 * h + 1] = (uint32_t)(uint16_t)r[4 * h + 2] | ((uint32_t)(uint16_t)r[4 * h + 3] << 16);
            } else {
                d[h] = (uint32_t)(uint8_t)r[4 * h] | ((uint32_t)(uint8_t)r[4 * h + 1] << 8) |
                       ((uint32_t)(uint8_t)r[4 * h + 2] << 16) | ((uint32_t)(uint8_t)r[4 * h + 3] << 24);
            }
        }
    }
}

template <typename TS>
__host__ __device__ __forceinline__ void int_contig_writeout(const IntParams &p, const TS *otile, int tid,
                                                             uint32_t bid)
{
    constexpr int TR = kIntContigRows;
    constexpr int A = 16 / (int)sizeof(TS);
    const IntContigTile t(p, bid);
    TS *out = (TS *)p.out;
    constexpr int SX = kIntContigOut;
    const int PO = p.opitch;
    if (p.vec) {
        constexpr int VR = SX / A;
        for (int v = tid; v < TR * VR; v += kIntThreads) {
            const int r = v / VR, cv = v - r * VR;
            const int64_t row = t.row0 + r, x = t.x0 + cv * A;
            if (row >= p.outer || x >= p.n) continue;
            const uint32_t *sw = reinterpret_cast<const uint32_t *>(otile + r * PO + cv * A);
            uint4 val = make_uint4(sw[0], sw[1], sw[2], sw[3]);
            TS *dst = out + row * p.n + x;
            if (p.epilogue != B200_EPI_STORE) {
                TS el[A], prev[A];
                memcpy(el, &val, 16);
                if (epilogue_reads_out(p.epilogue)) {
                    const uint4 pv = *reinterpret_cast<const uint4 *>(dst);
                    memcpy(prev, &pv, 16);
                } else {
#pragma unroll
                    for (int e = 0; e < A; ++e) prev[e] = TS(0);
                }
#pragma unroll
                for (int e = 0; e < A; ++e) el[e] = apply_epilogue<TS>(p.epilogue, prev[e], el[e]);
                memcpy(&val, el, 16);
            }
            *reinterpret_cast<uint4 *>(dst) = val;
        }
    } else {
        for (int e = tid; e < TR * SX; e += kIntThreads) {
            const int r = e / SX, cc = e - r * SX;
            const int64_t row = t.row0 + r, x = t.x0 + cc;
            if (row >= p.outer || x >= p.n) continue;
            TS *dst = out + row * p.n + x;
            TS v = otile[r * PO + cc];
            if (p.epilogue != B200_EPI_STORE) {
                const TS prev = epilogue_reads_out(p.epilogue) ? *dst : TS(0);
                v = apply_epilogue<TS>(p.epilogue, prev, v);
            }
            *dst = v;
        }
    }
}

// ---------------------------------------------------------------------------
// host: eligibility and launch geometry.  Returns 0 when the call is eligible
// and `p`, `contig`, `kind`, `smem_bytes`, `blocks` are filled in, -1 otherwise.
// ---------------------------------------------------------------------------
struct IntPlan {
    IntParams p;
    int contig, kind;
    size_t smem_bytes;
    int64_t blocks;
};

inline int plan_corr1d_int(const Corr1dArgs &a, IntPlan &pl, int symmetry)
{
    const int dt = a.dtype;
    if (dt != B200_U8 && dt != B200_I8 && dt != B200_U16 && dt != B200_I16) return -1;
    if (a.nw < 1 || a.nw > kMaxTapsInt) return -1;
    if (a.peer_lo || a.peer_hi) return -1;
    for (int i = 0; i < a.nw; ++i)
        if (!std::isfinite(a.w[i]) || std::fabs(a.w[i]) >= 0x1p900) return -1;
    const int sz = dtype_size(dt);
    const int A = 16 / sz;
    IntParams &p = pl.p;
    p.cval = 0;
    p.hi_word = kIntHigh;
    if (a.mode == B200_MODE_CONSTANT) {
        // the kernels hold cval as an array element; anything else keeps the double path
        const double lo_v = dt == B200_I8 ? -128.0 : (dt == B200_I16 ? -32768.0 : 0.0);
        const double hi_v = dt == B200_U8 ? 255.0 : (dt == B200_I8 ? 127.0 : (dt == B200_U16 ? 65535.0 : 32767.0));
        if (!(a.cval >= lo_v && a.cval <= hi_v) || a.cval != std::floor(a.cval)) return -1;
        p.cval = (int)a.cval;
    }
    const int L = a.nw;
    p.in = a.in; p.out = a.out;
    p.outer = a.g.outer; p.n = a.g.n; p.inner = a.g.inner;
    p.pad_lo = a.pad_lo; p.pad_hi = a.pad_hi;
    p.nw = L; p.lo = L / 2 + a.origin; p.mode = a.mode; p.epilogue = a.epilogue;
    pl.kind = symmetry;
    if (symmetry != 0) {
        p.nsteps = L / 2;
        p.centre = L / 2;
    } else {
        p.nsteps = L - 1;
        p.centre = L - 1;
    }
    for (int i = 0; i < L; ++i) {
        p.w[i] = a.w[i];
        p.c[i] = -(0x1p52 * a.w[i]);
    }
    pl.contig = a.g.inner == 1;
    if (pl.contig) {
        if (a.pad_lo || a.pad_hi) return -1;
        if (p.n >= (1ll << 31) - 4096) return -1;
        p.vec = (p.n % A == 0) && ((uintptr_t)a.in % 16 == 0) && ((uintptr_t)a.out % 16 == 0);
        p.tile_out = kIntContigOut;
        p.shift = (A - p.lo % A) % A;
        p.tile_len = (p.shift + p.tile_out + L - 1 + A - 1) / A * A;
        p.pitch = p.tile_len + 4 / sz;
        p.opitch = p.tile_out + 4 / sz;
        p.vpr_magic = (uint32_t)(((1ull << 32) + (p.tile_len / A) - 1) / (p.tile_len / A));
        p.c_tiles = (int)((p.n + p.tile_out - 1) / p.tile_out);
        p.n_tiles = 1;
        const int64_t row_tiles = (p.outer + kIntContigRows - 1) / kIntContigRows;
        pl.blocks = row_tiles * p.c_tiles;
        pl.smem_bytes = kIntTapBytes + (size_t)kIntContigRows * (p.pitch + p.opitch) * sz;
    } else {
        if (p.inner >= (1ll << 31) || p.n >= (1ll << 31) - 4096) return -1;
        p.vec = (p.inner % A == 0) && ((uintptr_t)a.in % 16 == 0);
        p.tile_out = 32;
        p.shift = 0;
        p.tile_len = p.tile_out + L - 1;
        p.pitch = kIntStridedW; p.opitch = 0;
        p.n_tiles = (int)((p.n + p.tile_out - 1) / p.tile_out);
        p.c_tiles = (int)((p.inner + kIntStridedW - 1) / kIntStridedW);
        pl.blocks = p.outer * (int64_t)p.n_tiles * p.c_tiles;
        pl.smem_bytes = kIntTapBytes + kIntRowTableBytes + (size_t)p.tile_len * kIntStridedW * sz;
    }
    if (pl.blocks < 1 || pl.blocks >= (1ll << 31)) return -1;
    return 0;
}

}  // namespace b200
