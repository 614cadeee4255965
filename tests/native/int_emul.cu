// int_emul.cu -- CPU execution of the exact tiled integer correlate1d kernels.
//
// Test infrastructure: corr1d_int.cuh keeps every phase of the two kernels as a
// __host__ __device__ function of (thread, block), so this file runs the very
// code of the CUDA kernels on the host -- block by block, phase by phase (a
// __syncthreads() in the kernel is the boundary between two loops over the
// threads here) -- and the CPU test suite compares the result with the oracle.
// Built by tests/test_int_emul.py with `nvcc -x cu` (host code only; no GPU).
#include "../../dask_image_b200/csrc/corrnd_int.cuh"
#include <vector>
#include <cstdlib>

namespace b200 {
// the launchers' diagnostics are not linked into this harness
int set_error(int code, const char *, ...) { return code; }
}  // namespace b200

using namespace b200;

// NQ = outputs per lane and walk: 4 is the shipped kernel; 8 (B200_INT_NQ=8) exercises the generic form
// of the walk (measured slower on the B200, profiles/r2_int_ab.log, and not built into the library)
template <typename TS, int KIND, int NQ>
static void run_blocks(const IntPlan &pl)
{
    const IntParams &p = pl.p;
    std::vector<unsigned char> smem(pl.smem_bytes + 16);
    IntTap *wc = reinterpret_cast<IntTap *>(smem.data());
    TS *tile = reinterpret_cast<TS *>(smem.data() + kIntTapBytes);
    for (uint32_t bid = 0; bid < (uint32_t)pl.blocks; ++bid) {
        // poison the tile so that a read of an element no thread staged shows up
        for (size_t i = 0; i < smem.size(); ++i) smem[i] = 0xA5;
        for (int t = 0; t < kIntThreads; ++t) int_stage_taps(p, wc, t);
        if (pl.contig) {
            TS *otile = tile + kIntContigRows * p.pitch;
            for (int t = 0; t < kIntThreads; ++t) int_contig_stage<TS>(p, tile, t, bid);
            for (int t = 0; t < kIntThreads; ++t) int_contig_compute<TS, KIND, NQ>(p, tile, otile, wc, t, bid);
            for (int t = 0; t < kIntThreads; ++t) int_contig_writeout<TS>(p, otile, t, bid);
        } else {
            int64_t *rows = reinterpret_cast<int64_t *>(smem.data() + kIntTapBytes);
            TS *stile = reinterpret_cast<TS *>(smem.data() + kIntTapBytes + kIntRowTableBytes);
            for (int t = 0; t < kIntThreads; ++t) int_strided_stage_rows(p, rows, t, bid);
            for (int t = 0; t < kIntThreads; ++t) int_strided_stage<TS>(p, stile, rows, t, bid);
            for (int t = 0; t < kIntThreads; ++t) int_strided_compute<TS, KIND, NQ>(p, stile, wc, t, bid);
        }
    }
}

template <typename TS, int NQ>
static void run_kind_nq(const IntPlan &pl)
{
    if (pl.kind > 0) run_blocks<TS, 1, NQ>(pl);
    else if (pl.kind < 0) run_blocks<TS, -1, NQ>(pl);
    else run_blocks<TS, 0, NQ>(pl);
}

template <typename TS>
static void run_kind(const IntPlan &pl)
{
    const char *e = getenv("B200_INT_NQ");
    if (e && atoi(e) == 8) run_kind_nq<TS, 8>(pl);
    else run_kind_nq<TS, 4>(pl);
}

// Same argument meaning as b200_correlate1d (host pointers).  Returns 0 when the
// tiled kernels take the call, -1 when they decline it (the library then runs
// the generic kernel), and reports which kernel and geometry through `info`:
// {contig, kind, blocks, smem_bytes, vec}.
extern "C" int emul_correlate1d_int(const void *in, void *out, int dtype, int ndim, const int64_t *shape, int axis,
                                    const double *weights, int64_t nw, int mode, double cval, int64_t origin,
                                    int64_t pad_lo, int64_t pad_hi, int epilogue, int64_t *info)
{
    Corr1dArgs a;
    a.in = in; a.out = out; a.dtype = dtype;
    a.g = line_geometry(ndim, shape, axis);
    a.w = weights; a.nw = (int)nw; a.mode = mode; a.cval = cval; a.origin = (int)origin;
    a.pad_lo = pad_lo; a.pad_hi = pad_hi; a.epilogue = epilogue; a.flags = 0; a.stream = nullptr;
    IntPlan pl{};
    if (plan_corr1d_int(a, pl, classify_symmetry(a.w, a.nw)) != 0) return -1;
    if (pl.smem_bytes > 48 * 1024) return -1;
    info[0] = pl.contig; info[1] = pl.kind; info[2] = pl.blocks; info[3] = (int64_t)pl.smem_bytes; info[4] = pl.p.vec;
    info[5] = 0; info[6] = 0; info[7] = 0;
    switch (dtype) {
    case B200_U8: run_kind<uint8_t>(pl); break;
    case B200_I8: run_kind<int8_t>(pl); break;
    case B200_U16: run_kind<uint16_t>(pl); break;
    default: run_kind<int16_t>(pl); break;
    }
    return 0;
}

template <typename TS>
static void run_nd_blocks(const IntNdPlan &pl)
{
    const IntNdParams &p = pl.p;
    std::vector<unsigned char> smem(pl.smem_bytes + 16);
    IntTap *wc = reinterpret_cast<IntTap *>(smem.data());
    int64_t *table = reinterpret_cast<int64_t *>(smem.data() + int_nd_tap_bytes(p.ntaps));
    TS *tile = reinterpret_cast<TS *>(smem.data() + int_nd_tap_bytes(p.ntaps) + int_nd_table_bytes(p));
    TS *otile = tile + p.KZ * p.stage_lines * p.pitch;
    for (uint32_t bid = 0; bid < (uint32_t)pl.blocks; ++bid) {
        for (size_t i = 0; i < smem.size(); ++i) smem[i] = 0xA5;
        for (int t = 0; t < kIntThreads; ++t) int_nd_stage_taps(p, wc, t);
        for (int t = 0; t < kIntThreads; ++t) int_nd_stage_lines(p, table, t, bid);
        for (int t = 0; t < kIntThreads; ++t) int_nd_stage<TS>(p, tile, table, t, bid);
        for (int t = 0; t < kIntThreads; ++t) int_nd_compute<TS>(p, tile, otile, wc, t, bid);
        for (int t = 0; t < kIntThreads; ++t) int_nd_writeout<TS>(p, otile, t, bid);
    }
}

// Same argument meaning as b200_correlate_nd (host pointers); info = {blocks, smem_bytes, vec, tile_out}.
extern "C" int emul_correlate_nd_int(const void *in, void *out, int dtype, int ndim, const int64_t *shape,
                                     const double *weights, const int64_t *wshape, const int64_t *origins, int mode,
                                     double cval, int64_t pad_lo, int64_t pad_hi, int64_t *info)
{
    CorrNdArgs a;
    a.in = in; a.out = out; a.dtype = dtype; a.ndim = ndim;
    for (int d = 0; d < ndim; ++d) {
        a.shape[d] = shape[d]; a.wshape[d] = wshape[d]; a.origins[d] = origins[d];
    }
    a.w = weights; a.mode = mode; a.cval = cval; a.pad_lo = pad_lo; a.pad_hi = pad_hi;
    a.flags = 0; a.stream = nullptr;
    IntNdPlan pl{};
    if (plan_corrnd_int(a, pl) != 0) return -1;
    pl.p.taps = pl.taps.data();
    info[0] = pl.blocks; info[1] = (int64_t)pl.smem_bytes; info[2] = pl.p.vec; info[3] = pl.p.tile_out;
    switch (dtype) {
    case B200_U8: run_nd_blocks<uint8_t>(pl); break;
    case B200_I8: run_nd_blocks<int8_t>(pl); break;
    case B200_U16: run_nd_blocks<uint16_t>(pl); break;
    default: run_nd_blocks<int16_t>(pl); break;
    }
    return 0;
}
