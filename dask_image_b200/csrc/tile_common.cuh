// tile_common.cuh -- device building blocks shared by the TMA-tiled kernels
// (corr1d_strided.cu / corr1d_contig.cu: 1-D passes and box filter; corrnd_tiled.cu: direct N-D
// correlate): 16-byte register vectors, packed-FFMA2 accumulators, storage
// widening loads, the contiguous-axis tap engine and wide global stores.
#pragma once
#include "common.cuh"
#include "tma.cuh"
#include <cstdlib>

namespace b200 {

constexpr size_t kSmemPerSm = 227 * 1024;
// Dynamic shared memory opt-in of every tiled kernel: always the device maximum, never the size of
// the launch at hand -- the attribute is per function, and concurrent host threads launching the same
// instantiation with different tile sizes would otherwise lower it under each other's launches.
constexpr int kSmemOptIn = 227 * 1024 - 1152;   // less the kernels' static shared memory (patch table, barrier)

static inline int env_int(const char *name, int dflt)
{
    const char *s = getenv(name);
    return (s && *s) ? atoi(s) : dflt;
}

template <typename TS> constexpr int dtype_code();
template <> constexpr int dtype_code<float>() { return B200_F32; }
template <> constexpr int dtype_code<double>() { return B200_F64; }
template <> constexpr int dtype_code<uint8_t>() { return B200_U8; }
template <> constexpr int dtype_code<uint16_t>() { return B200_U16; }

template <typename T, int V>
struct alignas(16) Vec {
    T v[V];
};

// ---------------------------------------------------------------------------
// Accumulator blocks.  Acc<T, N> holds N accumulators and consumes 16-byte
// vectors.  The float specialisation keeps them as float2 pairs and issues the
// sm_100 packed FFMA2 (fma.rn.f32x2, broadcast weight from a uniform
// register): half the issue slots of scalar FFMA for the same math.
// ---------------------------------------------------------------------------
template <typename T, int V>
struct VecAcc {
    T a[V];
    __device__ __forceinline__ void zero()
    {
#pragma unroll
        for (int i = 0; i < V; ++i) a[i] = T(0);
    }
    __device__ __forceinline__ void fma(const Vec<T, V> &x, T w)
    {
#pragma unroll
        for (int i = 0; i < V; ++i) a[i] = ::fma(w, x.v[i], a[i]);
    }
    __device__ __forceinline__ T get(int i) const { return a[i]; }
};

template <int V>
struct VecAcc<float, V> {
    static_assert(V % 2 == 0, "packed pairs");
    float2 a[V / 2];
    __device__ __forceinline__ void zero()
    {
#pragma unroll
        for (int i = 0; i < V / 2; ++i) a[i] = make_float2(0.f, 0.f);
    }
    __device__ __forceinline__ void fma(const Vec<float, V> &x, float w)
    {
        const float2 w2 = make_float2(w, w);
#pragma unroll
        for (int i = 0; i < V / 2; ++i) a[i] = __ffma2_rn(make_float2(x.v[2 * i], x.v[2 * i + 1]), w2, a[i]);
    }
    __device__ __forceinline__ float get(int i) const { return (i & 1) ? a[i >> 1].y : a[i >> 1].x; }
};

// Running extreme of the taps instead of a weighted sum (minimum / maximum
// filters): same interface as VecAcc, the weight is ignored.  Selection does not
// round, so the result is exact for floats and for integers widened to float.
template <typename T, int V, bool IS_MAX>
struct VecExt {
    T a[V];
    __device__ __forceinline__ void zero()
    {
#pragma unroll
        for (int i = 0; i < V; ++i) a[i] = IS_MAX ? -INFINITY : INFINITY;
    }
    __device__ __forceinline__ void fma(const Vec<T, V> &x, T)
    {
#pragma unroll
        for (int i = 0; i < V; ++i) a[i] = IS_MAX ? fmaxf(a[i], x.v[i]) : fminf(a[i], x.v[i]);
    }
    __device__ __forceinline__ T get(int i) const { return a[i]; }
};

// OP: 0 = correlation (FMA), 1 = minimum, 2 = maximum
template <typename T, int V, int OP> struct TapAcc { using type = VecAcc<T, V>; };
template <int V> struct TapAcc<float, V, 1> { using type = VecExt<float, V, false>; };
template <int V> struct TapAcc<float, V, 2> { using type = VecExt<float, V, true>; };

// 16-byte aligned base of the dynamic shared memory (TMA destinations need
// 128-byte alignment; the launcher over-allocates by 128).  Pointer arithmetic
// on the __shared__ array keeps the shared address space (LDS, not LD).
template <typename T>
__device__ __forceinline__ T *aligned_smem(unsigned char *raw)
{
    const uint32_t pad = (128u - (smem_u32(raw) & 127u)) & 127u;
    return reinterpret_cast<T *>(raw + pad);
}

// ---------------------------------------------------------------------------
// Storage <-> arithmetic vectors.  A thread-level block is V elements: 16 bytes
// when TS == T, 4 elements (8 / 4 bytes) when TS is uint16_t / uint8_t.
// Integer elements are widened with one PRMT each (bits 0x4B000000 | v are the
// float 2^23 + v) and a packed subtract of 2^23 -- exact, no I2F.
// ---------------------------------------------------------------------------
template <typename T, typename TS>
struct StorageV {
    static constexpr int V = sizeof(TS) == sizeof(T) ? 16 / (int)sizeof(T) : 4;
};

// strided-axis passes: a thread's column vector is 16 bytes of floats, or 8
// integer elements (16 bytes of uint16, 8 bytes of uint8) widened to 8 floats
template <typename T, typename TS>
struct StridedV {
    static constexpr int V = sizeof(TS) == sizeof(T) ? 16 / (int)sizeof(T) : 8;
};

template <typename T, typename TS, int V>
__device__ __forceinline__ Vec<T, V> load_block(const TS *ptr)
{
    if constexpr (sizeof(TS) == sizeof(T)) {
        return *reinterpret_cast<const Vec<T, V> *>(ptr);
    } else if constexpr (V == 8) {
        static_assert(sizeof(T) == 4, "integer storage widens to floats");
        uint32_t f[8];
        if constexpr (sizeof(TS) == 2) {
            const uint4 raw = *reinterpret_cast<const uint4 *>(ptr);
            const uint32_t wds[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                f[2 * i] = __byte_perm(wds[i], 0x4B000000u, 0x7410);
                f[2 * i + 1] = __byte_perm(wds[i], 0x4B000000u, 0x7432);
            }
        } else {
            const uint2 raw = *reinterpret_cast<const uint2 *>(ptr);
            const uint32_t wds[2] = {raw.x, raw.y};
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                f[4 * i] = __byte_perm(wds[i], 0x4B000000u, 0x7440);
                f[4 * i + 1] = __byte_perm(wds[i], 0x4B000000u, 0x7441);
                f[4 * i + 2] = __byte_perm(wds[i], 0x4B000000u, 0x7442);
                f[4 * i + 3] = __byte_perm(wds[i], 0x4B000000u, 0x7443);
            }
        }
        const float2 bias = make_float2(-8388608.f, -8388608.f);
        Vec<T, V> r;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 a = __fadd2_rn(make_float2(__uint_as_float(f[2 * i]), __uint_as_float(f[2 * i + 1])), bias);
            r.v[2 * i] = a.x; r.v[2 * i + 1] = a.y;
        }
        return r;
    } else {
        static_assert(V == 4 && sizeof(T) == 4, "integer storage widens to 4 floats");
        uint32_t f[4];
        if constexpr (sizeof(TS) == 2) {
            const uint2 raw = *reinterpret_cast<const uint2 *>(ptr);
            f[0] = __byte_perm(raw.x, 0x4B000000u, 0x7410);
            f[1] = __byte_perm(raw.x, 0x4B000000u, 0x7432);
            f[2] = __byte_perm(raw.y, 0x4B000000u, 0x7410);
            f[3] = __byte_perm(raw.y, 0x4B000000u, 0x7432);
        } else {
            const uint32_t raw = *reinterpret_cast<const uint32_t *>(ptr);
            f[0] = __byte_perm(raw, 0x4B000000u, 0x7440);
            f[1] = __byte_perm(raw, 0x4B000000u, 0x7441);
            f[2] = __byte_perm(raw, 0x4B000000u, 0x7442);
            f[3] = __byte_perm(raw, 0x4B000000u, 0x7443);
        }
        const float2 bias = make_float2(-8388608.f, -8388608.f);
        const float2 a = __fadd2_rn(make_float2(__uint_as_float(f[0]), __uint_as_float(f[1])), bias);
        const float2 b = __fadd2_rn(make_float2(__uint_as_float(f[2]), __uint_as_float(f[3])), bias);
        Vec<T, V> r;
        r.v[0] = a.x; r.v[1] = a.y; r.v[2] = b.x; r.v[3] = b.y;
        return r;
    }
}

// Integer box filter epilogue.  s is the exact window sum (an integer below
// 2^23 held in a float); the result is trunc(s / size) as the reference
// computes it (double division, C cast).  With s2 = s + 0.5 - size/2,
//     fma(s2, 1/size, 1.5 * 2^23)
// rounds (s + 0.5)/size - 0.5 to the nearest integer in ONE rounding, which is
// floor(s/size); the low bits of the float's bit pattern are that integer.
// The launcher proves the identity for every reachable sum before it selects
// this kernel (uniform_int_exact()).
template <typename T>
__device__ __forceinline__ uint32_t box_quotient_bits(T s, T qinv, T qc0)
{
    return __float_as_uint(fmaf((float)s + (float)qc0, (float)qinv, 12582912.f));
}

// N quotients (the low bits of q[i]) packed to the storage type and written with one store
template <typename TS, int N>
__device__ __forceinline__ void pack_box_quotients(TS *dst, const uint32_t (&q)[N])
{
    static_assert(N == 4 || N == 8, "4 or 8 outputs per store");
    if constexpr (sizeof(TS) == 2) {
        uint32_t w[N / 2];
#pragma unroll
        for (int i = 0; i < N / 2; ++i) w[i] = __byte_perm(q[2 * i], q[2 * i + 1], 0x5410);
        if constexpr (N == 4) *reinterpret_cast<uint2 *>(dst) = make_uint2(w[0], w[1]);
        else *reinterpret_cast<uint4 *>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
    } else {
        uint32_t w[N / 4];
#pragma unroll
        for (int i = 0; i < N / 4; ++i) {
            const uint32_t lo = __byte_perm(q[4 * i], q[4 * i + 1], 0x0040);
            const uint32_t hi = __byte_perm(q[4 * i + 2], q[4 * i + 3], 0x0040);
            w[i] = __byte_perm(lo, hi, 0x5410);
        }
        if constexpr (N == 4) *reinterpret_cast<uint32_t *>(dst) = w[0];
        else *reinterpret_cast<uint2 *>(dst) = make_uint2(w[0], w[1]);
    }
}

template <typename T, typename TS, int N>
__device__ __forceinline__ void store_box_quotients(TS *dst, const T (&s)[N], T qinv, T qc0)
{
    uint32_t q[N];
#pragma unroll
    for (int i = 0; i < N; ++i) q[i] = box_quotient_bits<T>(s[i], qinv, qc0);
    pack_box_quotients<TS, N>(dst, q);
}

// 32-byte global stores/loads (sm_100 STG.E.256 / LDG.E.256): a contiguous-axis
// thread owns 8 consecutive floats (4 doubles) of a row; one 256-bit store
// writes whole 32-byte sectors instead of two half-sector 128-bit stores.
__device__ __forceinline__ void st_global_256(float *dst, const float (&v)[8])
{
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst), "f"(v[0]), "f"(v[1]), "f"(v[2]),
                 "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
                 : "memory");
}
__device__ __forceinline__ void st_global_256(double *dst, const double (&v)[4])
{
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(dst), "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3])
                 : "memory");
}
__device__ __forceinline__ void ld_global_256(const float *src, float (&v)[8])
{
    asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                 : "l"(src));
}
__device__ __forceinline__ void ld_global_256(const double *src, double (&v)[4])
{
    asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(src));
}

// Per-thread state: R = 2V consecutive outputs of one row, three rotating
// 16-byte register blocks B (12 floats / 6 doubles of the row).
//
// double: scalar FMAs.  float: packed FFMA2 on aligned register pairs.  A tap
// at even data offset t pairs outputs (0,1)(2,3)(4,5)(6,7) with the aligned
// data pairs {v[j+t], v[j+t+1]}; a tap at odd offset would need misaligned
// pairs, so odd taps accumulate into a second set paired at ODD outputs
// (-1,0)(1,2)(3,4)(5,6)(7,8) -- aligned again -- and the two sets are summed at
// the end (the two half-wasted end pairs cost 1 extra FFMA2 per odd tap).
template <typename T, int V>
struct ContigAcc {
    static constexpr int R = 2 * V;
    T a[R];
    __device__ __forceinline__ void zero()
    {
#pragma unroll
        for (int j = 0; j < R; ++j) a[j] = T(0);
    }
    // tap at static data offset TT (0..V-1) inside the current group; U = group slot
    template <int U, int TT>
    __device__ __forceinline__ void tap(const Vec<T, V> (&B)[3], T w)
    {
#pragma unroll
        for (int j = 0; j < R; ++j) {
            const int m = j + TT;
            a[j] = ::fma(w, B[(U + m / V) % 3].v[m % V], a[j]);
        }
    }
    __device__ __forceinline__ T get(int j) const { return a[j]; }
};

template <>
struct ContigAcc<float, 4> {
    static constexpr int R = 8;
    float2 e[4];   // outputs (0,1)(2,3)(4,5)(6,7): even data offsets
    float2 o[5];   // outputs (-1,0)(1,2)(3,4)(5,6)(7,8): odd data offsets
    __device__ __forceinline__ void zero()
    {
#pragma unroll
        for (int i = 0; i < 4; ++i) e[i] = make_float2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 5; ++i) o[i] = make_float2(0.f, 0.f);
    }
    template <int U, int TT>
    __device__ __forceinline__ void tap(const Vec<float, 4> (&B)[3], float w)
    {
        const float2 w2 = make_float2(w, w);
        if constexpr ((TT & 1) == 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int m = 2 * i + TT;   // even
                const Vec<float, 4> &b = B[(U + m / 4) % 3];
                e[i] = __ffma2_rn(make_float2(b.v[m % 4], b.v[m % 4 + 1]), w2, e[i]);
            }
        } else {
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                const int m = 2 * i - 1 + TT;   // even, 0..10
                const Vec<float, 4> &b = B[(U + m / 4) % 3];
                o[i] = __ffma2_rn(make_float2(b.v[m % 4], b.v[m % 4 + 1]), w2, o[i]);
            }
        }
    }
    __device__ __forceinline__ float get(int j) const
    {
        const float ev = (j & 1) ? e[j >> 1].y : e[j >> 1].x;
        const float od = (j & 1) ? o[(j + 1) >> 1].x : o[j >> 1].y;
        return ev + od;
    }
};

// One group of V taps starting at tap index kb (a multiple of V): loads the
// 16-byte block two ahead into rotating slot (U+2)%3 and applies taps
// [T0, T1) of the group.  U = (kb / V) % 3 is static so the register rotation
// is free; T0/T1 are static so partial groups have no per-tap tests.
template <typename T, typename TS, int V, int U, int T0, int T1, typename ACC>
__device__ __forceinline__ void contig_group(Vec<T, V> (&B)[3], ACC &acc, const TS *srow, const T *w, int kb)
{
    B[(U + 2) % 3] = load_block<T, TS, V>(srow + kb + 2 * V);
    auto one = [&]<int TT>() {
        if constexpr (TT >= T0 && TT < T1) acc.template tap<U, TT>(B, w[kb + TT]);
    };
    one.template operator()<0>();
    one.template operator()<1>();
    if constexpr (V == 4) {
        one.template operator()<2>();
        one.template operator()<3>();
    }
}

// All taps of one tile row for a thread's R = 2V outputs.  `srow` points at the
// thread's first input element of the row (16-byte aligned), `w` at the row's
// L taps (the first SH of them are alignment padding and are skipped), L - REM
// is a multiple of V.  Control flow is warp-uniform.
template <typename T, typename TS, int V, int SH, int REM, typename ACC>
__device__ __forceinline__ void row_taps(ACC &acc, const TS *srow, const T *w, int L)
{
    Vec<T, V> B[3];
    B[0] = load_block<T, TS, V>(srow);
    B[1] = load_block<T, TS, V>(srow + V);
    const int Lfull = L - REM;
    int k = 0;
    if (Lfull >= V) {
        contig_group<T, TS, V, 0, SH, V>(B, acc, srow, w, 0);
        k = V;
        for (; k + 3 * V <= Lfull; k += 3 * V) {
            contig_group<T, TS, V, 1, 0, V>(B, acc, srow, w, k);
            contig_group<T, TS, V, 2, 0, V>(B, acc, srow, w, k + V);
            contig_group<T, TS, V, 0, 0, V>(B, acc, srow, w, k + 2 * V);
        }
        // 0, 1 or 2 full groups left, then the partial one
        if (k + V <= Lfull) {
            contig_group<T, TS, V, 1, 0, V>(B, acc, srow, w, k);
            k += V;
            if (k + V <= Lfull) {
                contig_group<T, TS, V, 2, 0, V>(B, acc, srow, w, k);
                k += V;
                if constexpr (REM > 0) contig_group<T, TS, V, 0, 0, REM>(B, acc, srow, w, k);
            } else {
                if constexpr (REM > 0) contig_group<T, TS, V, 2, 0, REM>(B, acc, srow, w, k);
            }
        } else {
            if constexpr (REM > 0) contig_group<T, TS, V, 1, 0, REM>(B, acc, srow, w, k);
        }
    } else {
        // fewer than V taps in total: a single partial group (REM > SH here)
        if constexpr (REM > 0) contig_group<T, TS, V, 0, SH, REM>(B, acc, srow, w, 0);
    }
}

}  // namespace b200
