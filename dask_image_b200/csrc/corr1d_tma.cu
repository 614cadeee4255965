// corr1d_tma.cu -- kernel selection for the TMA-tiled 1-D passes (sm_100a): the
// tensor-map encoder, the choice between the strided-axis unit (corr1d_strided.cu)
// and the contiguous-axis unit (corr1d_contig.cu) for the float passes
// (scipy/ndimage/_filters.py:852-858: one launch per 1-D pass of scipy's chain),
// the float and the exact uint8 / uint16 box filter (with the host proof of the
// fused-rounding quotient), and the tiled minimum / maximum filter.
// Accumulation is in the array's own float type (fp32 for f32, fp64 for f64)
// with FMA; B200_FLAG_EXACT selects generic.cu.  The integer box filter sums in
// fp32 (exact below 2^23) and divides with one proven-exact fused rounding
// (box_quotient_bits).
#include "corr1d_tma.cuh"

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

int launch_corr1d_tma(const Corr1dArgs &a)
{
    // NB the choice of kernel family must not depend on the number of planes
    // of the call: slabs of one volume (any GPU count) have to take the same
    // arithmetic path so that the whole-volume result is partition-invariant.
    const BoxInfo none;
    if (a.dtype != B200_F32 && a.dtype != B200_F64) return -1;
    return a.g.inner == 1 ? tma_contig_corr(a) : tma_strided_corr(a, none);
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
bool uniform_box_ok(int dtype, int size, int mode, double cval)
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

static int launch_uniform_box(const Uniform1dArgs &u)
{
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
    return u.g.inner == 1 ? tma_contig_box(a, box) : tma_strided_corr(a, box);
}

int launch_uniform1d_fast(const Uniform1dArgs &u)
{
    if (u.size < 2) return -1;
    if (u.dtype == B200_U16 || u.dtype == B200_U8) return launch_uniform_box(u);
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
    return tma_peer_box(a, box);
}

// ---------------------------------------------------------------------------
// minimum / maximum filter on the tiled kernels
// ---------------------------------------------------------------------------
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
    if (u.g.inner == 1) return tma_contig_minmax(a, u.is_max != 0);
    return tma_strided_minmax(a, box, u.is_max != 0);
}

// minimum / maximum filter on the tiled kernels: float32 and uint8 / uint16
// (exact in fp32).  The tile holds storage-type values, so a constant-mode cval
// must be one; -1 = not eligible (the generic exact kernel takes the call).
int launch_minmax1d_fast(const MinMax1dArgs &u)
{
    if (u.size < 2) return -1;
    if (u.dtype == B200_F32) {
        if (u.mode == B200_MODE_CONSTANT && (double)(float)u.cval != u.cval) return -1;
        return launch_minmax_typed(u);
    }
    if (u.dtype == B200_U16 || u.dtype == B200_U8) {
        const double vmax = u.dtype == B200_U16 ? 65535.0 : 255.0;
        if (u.mode == B200_MODE_CONSTANT &&
            (!(u.cval >= 0.0 && u.cval <= vmax) || u.cval != (double)(int64_t)u.cval))
            return -1;
        return launch_minmax_typed(u);
    }
    return -1;
}

}  // namespace b200
