// corr1d_tma.cu -- TMA-staged shared-memory 1-D correlation passes (sm_100a).
//
// The two kernels here are the float hot path behind gaussian_filter and its
// derivative / gradient-magnitude / laplace variants (one launch per 1-D pass
// of scipy's chain, scipy/ndimage/_filters.py:852-858):
//
//  * corr1d_strided_tma_kernel  -- filtered axis is NOT the contiguous one.
//    The array is viewed as [outer][n][inner]; one CTA stages a
//    (tile_rows + taps - 1) x W box (rows along the filtered axis, W contiguous
//    elements) with ONE cp.async.bulk.tensor.3d, then every thread owns a 16-byte
//    column vector and R consecutive output rows, sliding a register window of
//    R rows down the tile: per tap 1 LDS.128 + R*V FMAs, nothing else.
//
//  * corr1d_contig_tma_kernel   -- filtered axis IS the contiguous one.  View
//    [rows][n]; a CTA stages a TR x P box with one cp.async.bulk.tensor.2d where
//    the pitch P is an odd multiple of 16 bytes, so that the 8 lanes of a
//    quarter warp (4 lanes x 32 bytes along x, 2 rows) hit 8 distinct 16-byte
//    bank groups: conflict-free LDS.128 straight out of the dense TMA layout.
//    Every thread produces R = 2V consecutive outputs of one row from three
//    rotating 16-byte register blocks: per V taps 1 LDS.128 + R*V FMAs.
//
// Out-of-range rows/columns are zero-filled by TMA; CTAs on an array edge then
// patch exactly those smem rows/columns through the closed-form extension map
// (reflect / mirror / nearest / wrap / constant), so interior CTAs carry no
// boundary logic at all.  Accumulation is in the array's own float type
// (fp32 for f32, fp64 for f64) with FMA; B200_FLAG_EXACT selects generic.cu.
#include "common.cuh"
#include "tma.cuh"
#include <mutex>
#include <cstdlib>

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
constexpr int kMaxTapsTma = 200;

template <typename T>
struct Corr1dTmaParams {
    const T *in;   // plane/row 0 of the array (edge patch-up loads)
    T *out;
    int64_t outer, n, inner;
    int64_t pad_lo, pad_hi;
    int nw, lo, mode, epilogue;
    int tile_rows;   // strided: outputs per CTA along the filtered axis
    int smem_rows;   // strided: tile_rows + nw - 1
    int nb;          // strided: row blocks per thread
    int pitch;       // contig: smem row pitch in elements (V * odd)
    int n_tiles, c_tiles;
    T cval;
    T w[kMaxTapsTma];
};

template <typename T, int V>
struct alignas(16) Vec {
    T v[V];
};

template <typename T> __device__ __forceinline__ T fma_t(T a, T b, T c);
template <> __device__ __forceinline__ float fma_t<float>(float a, float b, float c) { return fmaf(a, b, c); }
template <> __device__ __forceinline__ double fma_t<double>(double a, double b, double c) { return fma(a, b, c); }

// ===========================================================================
// strided-axis pass
// ===========================================================================
template <typename T, int W, int TY, int R>
__global__ void __launch_bounds__((W / (16 / (int)sizeof(T))) * TY, 2)
corr1d_strided_tma_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ Corr1dTmaParams<T> p)
{
    constexpr int V = 16 / (int)sizeof(T);
    constexpr int TX = W / V;
    constexpr int NTHREADS = TX * TY;
    using VecT = Vec<T, V>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // TMA destinations must be 128-byte aligned; the launcher over-allocates by 128
    T *tile = reinterpret_cast<T *>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
    __shared__ __align__(8) uint64_t bar;

    const int tid = threadIdx.x;
    int64_t b = blockIdx.x;
    const int tn = (int)(b % p.n_tiles);
    b /= p.n_tiles;
    const int tc = (int)(b % p.c_tiles);
    const int64_t o = b / p.c_tiles;
    const int64_t row0 = (int64_t)tn * p.tile_rows;  // first output row of this CTA
    const int64_t col0 = (int64_t)tc * W;
    const int64_t g0 = row0 - p.lo;                  // array row held by smem row 0

    if (tid == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
    }
    __syncthreads();
    if (tid == 0) {
        mbar_arrive_expect_tx(&bar, (uint32_t)(p.smem_rows * W * (int)sizeof(T)));
        tma_load_3d(tile, &tmap, &bar, (int)col0, (int)(g0 + p.pad_lo), (int)o);
    }
    mbar_wait(&bar, 0);

    // rows of the tile that lie outside the resident range: patch through the
    // extension map (TMA left zeros there, which is already right for
    // mode=constant, cval=0)
    const bool low_edge = g0 < -p.pad_lo;
    const bool high_edge = g0 + p.smem_rows > p.n + p.pad_hi;
    if ((low_edge || high_edge) && !(p.mode == B200_MODE_CONSTANT && p.cval == T(0))) {
        const int warp = tid >> 5, lane = tid & 31;
        for (int r = warp; r < p.smem_rows; r += NTHREADS / 32) {
            const int64_t g = g0 + r;
            if (g >= -p.pad_lo && g < p.n + p.pad_hi) continue;
            const int64_t src = ext_index_padded(g, p.n, p.mode, p.pad_lo, p.pad_hi);
            const T *srow = (src == B200_USE_CVAL) ? nullptr : p.in + ((o * p.n + src) * p.inner + col0);
            for (int c = lane; c < W; c += 32) {
                T v = p.cval;
                if (srow) v = (col0 + c < p.inner) ? srow[c] : T(0);
                tile[r * W + c] = v;
            }
        }
        __syncthreads();
    }

    const int tx = tid % TX, ty = tid / TX;
    const int L = p.nw;
    const bool reads_out = epilogue_reads_out(p.epilogue);
    for (int blk = 0; blk < p.nb; ++blk) {
        const int r0 = (blk * TY + ty) * R;
        if (row0 + r0 >= p.n) break;
        const T *sbase = tile + r0 * W + tx * V;
        VecT win[R];
        T acc[R][V];
#pragma unroll
        for (int j = 0; j < R; ++j)
#pragma unroll
            for (int v = 0; v < V; ++v) acc[j][v] = T(0);
#pragma unroll
        for (int j = 0; j < R - 1; ++j) win[j] = *reinterpret_cast<const VecT *>(sbase + j * W);
        int k = 0;
        for (; k + R <= L; k += R) {
#pragma unroll
            for (int u = 0; u < R; ++u) {
                win[(u + R - 1) % R] = *reinterpret_cast<const VecT *>(sbase + (k + u + R - 1) * W);
                const T wk = p.w[k + u];
#pragma unroll
                for (int j = 0; j < R; ++j)
#pragma unroll
                    for (int v = 0; v < V; ++v) acc[j][v] = fma_t<T>(wk, win[(u + j) % R].v[v], acc[j][v]);
            }
        }
#pragma unroll
        for (int u = 0; u < R; ++u) {
            if (k + u < L) {
                win[(u + R - 1) % R] = *reinterpret_cast<const VecT *>(sbase + (k + u + R - 1) * W);
                const T wk = p.w[k + u];
#pragma unroll
                for (int j = 0; j < R; ++j)
#pragma unroll
                    for (int v = 0; v < V; ++v) acc[j][v] = fma_t<T>(wk, win[(u + j) % R].v[v], acc[j][v]);
            }
        }
        const int64_t col = col0 + tx * V;
        if (col < p.inner) {
#pragma unroll
            for (int j = 0; j < R; ++j) {
                const int64_t row = row0 + r0 + j;
                if (row < p.n) {
                    T *dst = p.out + ((o * p.n + row) * p.inner + col);
                    VecT res;
                    if (p.epilogue == B200_EPI_STORE) {
#pragma unroll
                        for (int v = 0; v < V; ++v) res.v[v] = acc[j][v];
                    } else {
                        VecT prev;
                        if (reads_out) prev = *reinterpret_cast<const VecT *>(dst);
#pragma unroll
                        for (int v = 0; v < V; ++v)
                            res.v[v] = apply_epilogue<T>(p.epilogue, reads_out ? prev.v[v] : T(0), acc[j][v]);
                    }
                    *reinterpret_cast<VecT *>(dst) = res;
                }
            }
        }
    }
}

// ===========================================================================
// contiguous-axis pass
// ===========================================================================
template <typename T, int TR>
__global__ void __launch_bounds__(256, 2)
corr1d_contig_tma_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ Corr1dTmaParams<T> p)
{
    constexpr int V = 16 / (int)sizeof(T);
    constexpr int R = 2 * V;        // outputs per thread
    constexpr int PC = 4 * R;       // patch columns (4 lanes along x)
    constexpr int SX = 4 * PC;      // tile columns (128 floats / 64 doubles = 512 B)
    constexpr int NPATCH = (TR / 8) * 4;
    using VecT = Vec<T, V>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // TMA destinations must be 128-byte aligned; the launcher over-allocates by 128
    T *tile = reinterpret_cast<T *>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
    __shared__ __align__(8) uint64_t bar;

    const int tid = threadIdx.x;
    const int P = p.pitch;
    int64_t b = blockIdx.x;
    const int tcx = (int)(b % p.c_tiles);      // tile index along the filtered (contiguous) axis
    const int64_t trow = b / p.c_tiles;
    const int64_t x0t = (int64_t)tcx * SX;     // first output column of this CTA
    const int64_t row0 = trow * TR;
    const int64_t g0 = x0t - p.lo;             // array column held by smem column 0
    const int64_t nrows = p.outer;

    if (tid == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
    }
    __syncthreads();
    if (tid == 0) {
        mbar_arrive_expect_tx(&bar, (uint32_t)(TR * P * (int)sizeof(T)));
        tma_load_2d(tile, &tmap, &bar, (int)g0, (int)row0);
    }
    mbar_wait(&bar, 0);

    const bool low_edge = g0 < 0;
    const bool high_edge = g0 + P > p.n;
    if ((low_edge || high_edge) && !(p.mode == B200_MODE_CONSTANT && p.cval == T(0))) {
        // columns [0, -g0) and [n - g0, P) of every tile row
        const int nlow = low_edge ? (int)(-g0 < P ? -g0 : P) : 0;
        const int hstart = high_edge ? (int)((p.n - g0) > 0 ? (p.n - g0) : 0) : P;
        const int nhigh = P - hstart;
        const int per_row = nlow + nhigh;
        for (int idx = tid; idx < per_row * TR; idx += 256) {
            const int r = idx / per_row;
            int c = idx - r * per_row;
            c = c < nlow ? c : hstart + (c - nlow);
            const int64_t src = ext_index(g0 + c, p.n, p.mode);
            T v = p.cval;
            if (src >= 0) v = (row0 + r < nrows) ? p.in[(row0 + r) * p.n + src] : T(0);
            tile[r * P + c] = v;
        }
        __syncthreads();
    }

    const int warp = tid >> 5, lane = tid & 31;
    const int L = p.nw;
    const bool reads_out = epilogue_reads_out(p.epilogue);
    for (int q = warp; q < NPATCH; q += 8) {
        const int rg = q % (TR / 8), cg = q / (TR / 8);
        const int r = rg * 8 + (lane >> 2);
        const int xl = cg * PC + (lane & 3) * R;
        const int64_t row = row0 + r;
        const int64_t x = x0t + xl;
        const T *srow = tile + r * P + xl;
        VecT B[3];
        T acc[R];
#pragma unroll
        for (int j = 0; j < R; ++j) acc[j] = T(0);
        B[0] = *reinterpret_cast<const VecT *>(srow);
        B[1] = *reinterpret_cast<const VecT *>(srow + V);
        int k = 0;
        for (; k + 3 * V <= L; k += 3 * V) {
#pragma unroll
            for (int u = 0; u < 3; ++u) {
                B[(u + 2) % 3] = *reinterpret_cast<const VecT *>(srow + k + (u + 2) * V);
#pragma unroll
                for (int t = 0; t < V; ++t) {
                    const T wk = p.w[k + u * V + t];
#pragma unroll
                    for (int j = 0; j < R; ++j) {
                        const int m = j + t;
                        acc[j] = fma_t<T>(wk, B[(u + m / V) % 3].v[m % V], acc[j]);
                    }
                }
            }
        }
#pragma unroll
        for (int u = 0; u < 3; ++u) {
            if (k + u * V < L) {
                B[(u + 2) % 3] = *reinterpret_cast<const VecT *>(srow + k + (u + 2) * V);
#pragma unroll
                for (int t = 0; t < V; ++t) {
                    if (k + u * V + t < L) {
                        const T wk = p.w[k + u * V + t];
#pragma unroll
                        for (int j = 0; j < R; ++j) {
                            const int m = j + t;
                            acc[j] = fma_t<T>(wk, B[(u + m / V) % 3].v[m % V], acc[j]);
                        }
                    }
                }
            }
        }
        if (row < nrows) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int64_t xx = x + h * V;
                if (xx < p.n) {
                    T *dst = p.out + (row * p.n + xx);
                    VecT res;
                    if (p.epilogue == B200_EPI_STORE) {
#pragma unroll
                        for (int v = 0; v < V; ++v) res.v[v] = acc[h * V + v];
                    } else {
                        VecT prev;
                        if (reads_out) prev = *reinterpret_cast<const VecT *>(dst);
#pragma unroll
                        for (int v = 0; v < V; ++v)
                            res.v[v] = apply_epilogue<T>(p.epilogue, reads_out ? prev.v[v] : T(0), acc[h * V + v]);
                    }
                    *reinterpret_cast<VecT *>(dst) = res;
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------
static int env_int(const char *name, int dflt)
{
    const char *s = getenv(name);
    return (s && *s) ? atoi(s) : dflt;
}

template <typename T>
static int launch_strided(const Corr1dArgs &a)
{
    constexpr int V = 16 / (int)sizeof(T);
    constexpr int W = 32 * V;   // 512-byte tile rows: one warp spans a row with 16-byte vectors
    constexpr int TY = 8, R = 8;
    constexpr int ROWS_PER_BLK = TY * R;   // 64
    const int es = (int)sizeof(T);
    const LineGeom &g = a.g;
    const int L = a.nw;
    if (L > kMaxTapsTma || ROWS_PER_BLK + L - 1 > 256) return -1;
    if ((g.inner * es) % 16 != 0) return -1;
    const int64_t n_tot = a.pad_lo + g.n + a.pad_hi;
    if (g.inner >= (1ll << 31) || n_tot >= (1ll << 31) || g.outer >= (1ll << 31)) return -1;
    const char *base = (const char *)a.in - a.pad_lo * g.inner * es;
    if (((uintptr_t)base & 15) || ((uintptr_t)a.out & 15)) return -1;

    // row blocks per CTA: as tall as shared memory (2 CTAs/SM), the 256-row TMA
    // box limit and the need for >= ~4 waves of CTAs allow
    const int c_tiles = (int)((g.inner + W - 1) / W);
    int nb = env_int("B200_TMA_NB", 0);
    if (nb <= 0) {
        nb = 1;
        const int sms = device_sm_count();
        for (int cand = 2; cand <= 3; ++cand) {
            const int rows = cand * ROWS_PER_BLK + L - 1;
            if (rows > 256 || (size_t)rows * W * es > 110 * 1024) break;
            const int64_t ctas = ((g.n + cand * ROWS_PER_BLK - 1) / (cand * ROWS_PER_BLK)) * c_tiles * g.outer;
            if (ctas < 4ll * sms) break;
            nb = cand;
        }
    }
    while (nb > 1 && (nb * ROWS_PER_BLK + L - 1 > 256)) --nb;
    const int tile_rows = nb * ROWS_PER_BLK;
    const int smem_rows = tile_rows + L - 1;
    const int64_t n_tiles = (g.n + tile_rows - 1) / tile_rows;
    const int64_t grid = n_tiles * c_tiles * g.outer;
    if (grid >= (1ll << 31)) return -1;

    CUtensorMap tmap;
    uint64_t dims[3] = {(uint64_t)g.inner, (uint64_t)n_tot, (uint64_t)g.outer};
    uint64_t strides[2] = {(uint64_t)g.inner * es, (uint64_t)g.inner * es * (uint64_t)n_tot};
    uint32_t box[3] = {(uint32_t)W, (uint32_t)smem_rows, 1u};
    if (a.pad_lo || a.pad_hi) {
        // pads exist only for axis 0, where outer == 1
        if (g.outer != 1) return -1;
    }
    if (encode_tensor_map(&tmap, a.dtype, (void *)base, 3, dims, strides, box) != 0) return -1;

    Corr1dTmaParams<T> p;
    p.in = (const T *)a.in; p.out = (T *)a.out;
    p.outer = g.outer; p.n = g.n; p.inner = g.inner;
    p.pad_lo = a.pad_lo; p.pad_hi = a.pad_hi;
    p.nw = L; p.lo = L / 2 + a.origin; p.mode = a.mode; p.epilogue = a.epilogue;
    p.tile_rows = tile_rows; p.smem_rows = smem_rows; p.nb = nb; p.pitch = W;
    p.n_tiles = (int)n_tiles; p.c_tiles = c_tiles;
    p.cval = (T)a.cval;
    for (int i = 0; i < L; ++i) p.w[i] = (T)a.w[i];

    auto kern = corr1d_strided_tma_kernel<T, W, TY, R>;
    const size_t smem = (size_t)smem_rows * W * es + 128;
    B200_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)grid, (W / V) * TY, smem, a.stream>>>(tmap, p);
    B200_CUDA_TRY(cudaGetLastError());
    count_launch();
    note_kernel(sizeof(T) == 4 ? "corr1d_tma_strided_f32" : "corr1d_tma_strided_f64");
    return B200_OK;
}

template <typename T>
static int launch_contig(const Corr1dArgs &a)
{
    constexpr int V = 16 / (int)sizeof(T);
    constexpr int R = 2 * V, SX = 16 * R, TR = 32;
    const int es = (int)sizeof(T);
    const LineGeom &g = a.g;
    const int L = a.nw;
    if (a.pad_lo || a.pad_hi) return -1;
    if (L > kMaxTapsTma) return -1;
    // pitch: smallest odd multiple of 16 bytes covering SX + L + V - 1 elements
    int pitch = SX + L + V - 1;
    pitch = (pitch + V - 1) / V;        // in 16-byte units
    if (!(pitch & 1)) ++pitch;
    pitch *= V;
    if (pitch > 256) return -1;
    if ((g.n * es) % 16 != 0) return -1;
    if (g.n >= (1ll << 31) || g.outer >= (1ll << 31)) return -1;
    if (((uintptr_t)a.in & 15) || ((uintptr_t)a.out & 15)) return -1;
    const int64_t c_tiles = (g.n + SX - 1) / SX;
    const int64_t r_tiles = (g.outer + TR - 1) / TR;
    const int64_t grid = c_tiles * r_tiles;
    if (grid >= (1ll << 31)) return -1;

    CUtensorMap tmap;
    uint64_t dims[2] = {(uint64_t)g.n, (uint64_t)g.outer};
    uint64_t strides[1] = {(uint64_t)g.n * es};
    uint32_t box[2] = {(uint32_t)pitch, (uint32_t)TR};
    if (encode_tensor_map(&tmap, a.dtype, (void *)a.in, 2, dims, strides, box) != 0) return -1;

    Corr1dTmaParams<T> p;
    p.in = (const T *)a.in; p.out = (T *)a.out;
    p.outer = g.outer; p.n = g.n; p.inner = 1;
    p.pad_lo = 0; p.pad_hi = 0;
    p.nw = L; p.lo = L / 2 + a.origin; p.mode = a.mode; p.epilogue = a.epilogue;
    p.tile_rows = TR; p.smem_rows = TR; p.nb = 1; p.pitch = pitch;
    p.n_tiles = (int)r_tiles; p.c_tiles = (int)c_tiles;
    p.cval = (T)a.cval;
    for (int i = 0; i < L; ++i) p.w[i] = (T)a.w[i];

    auto kern = corr1d_contig_tma_kernel<T, TR>;
    const size_t smem = (size_t)TR * pitch * es + 128;
    B200_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)grid, 256, smem, a.stream>>>(tmap, p);
    B200_CUDA_TRY(cudaGetLastError());
    count_launch();
    note_kernel(sizeof(T) == 4 ? "corr1d_tma_contig_f32" : "corr1d_tma_contig_f64");
    return B200_OK;
}

int launch_corr1d_tma(const Corr1dArgs &a)
{
    const int64_t total = a.g.outer * a.g.n * a.g.inner;
    // tiny arrays: one generic launch is as good, and keeps odd shapes simple
    if (!(a.flags & B200_FLAG_FORCE_FAST) && total < (1 << 16)) return -1;
    if (a.dtype == B200_F32)
        return a.g.inner == 1 ? launch_contig<float>(a) : launch_strided<float>(a);
    if (a.dtype == B200_F64)
        return a.g.inner == 1 ? launch_contig<double>(a) : launch_strided<double>(a);
    return -1;
}

}  // namespace b200
