// corr1d_tma.cuh -- declarations shared by the three translation units of the
// TMA-tiled 1-D passes: corr1d_strided.cu (filtered axis not contiguous, incl.
// the multi-GPU peer-halo kernel), corr1d_contig.cu (filtered axis contiguous)
// and corr1d_tma.cu (tensor-map encoder, kernel selection, box-filter proof).
// One unit per kernel family keeps every unit small enough to be compiled in
// one piece -- the library's SASS is then a function of the source alone
// (nvcc's --split-compile is not reproducible) -- and the units build in parallel.
#pragma once
#include "tile_common.cuh"
#include <mutex>
#include <cstdlib>
#include <cstring>
#include <cmath>

namespace b200 {

int encode_tensor_map(CUtensorMap *map, int dtype, void *base, int rank, const uint64_t *dims,
                      const uint64_t *strides_bytes, const uint32_t *box);

// ---------------------------------------------------------------------------
constexpr int kMaxTapsTma = 208;

// T = arithmetic type (float / double), TS = storage type of the array: TS == T
// for the float passes; TS = uint8_t / uint16_t with T = float for the exact
// integer box filter (see the integer epilogue below).
template <typename T, typename TS = T>
struct Corr1dTmaParams {
    const TS *in;   // plane/row 0 of the array (edge patch-up loads)
    TS *out;
    int64_t outer, n, inner;
    int64_t pad_lo, pad_hi;
    int nw, lo, mode, epilogue;
    int tile_rows;   // strided: outputs per CTA tile along the filtered axis
    int smem_rows;   // strided: tile_rows + nw - 1
    int nb;          // strided: row blocks per thread
    int pitch;       // contig: smem row pitch in elements (V * odd)
    int shift;       // contig: leading taps to skip (box start aligned down to 16 bytes)
    int ncg;         // contig: 4*R-column groups per tile
    int wide_store;  // contig: rows and base are 32-byte aligned -> 256-bit stores
    int n_tiles, c_tiles;
    TS cval;
    T qinv, qc0;     // integer box filter: 1/size and the (0.5 - size/2) offset
    T w[kMaxTapsTma];
};

// ===========================================================================
// strided-axis pass: one tile per CTA, 3 CTAs resident per SM
// ===========================================================================
// REM = taps % R is a template parameter so that the tap tail is straight-line
// code; nothing in the tap loops is divergent, which lets ptxas keep the
// weights on the uniform datapath (LDCU -> FFMA2 with a UR operand).
// Peer halos (multi-GPU axis-0 pass, sharded.py): the planes below / above the
// slab are NOT copied next to it first -- the kernel reads them where they
// live, in the neighbour GPU's memory, with its own TMA loads over NVLink, so
// the "exchange" is part of the tile pipeline of the pass itself: no NCCL
// kernel competing for SMs, no staging planes, no separate edge launches.
// The first tile of the axis takes its top `lo` rows from `lo` (the previous
// rank's last planes) and the rest from a shorter local box; the last tile takes
// the `hi` rows behind the slab from `hi`.  Every other tile is unchanged.
struct PeerMaps {
    CUtensorMap first;   // local slab, box rows = smem_rows - lo (first tile, below its peer rows)
    CUtensorMap last;    // local slab, box rows = k_last         (last tile, above which peer rows follow)
    CUtensorMap lo;      // the lo planes that precede the slab (in a peer's memory)
    CUtensorMap hi;      // the hi planes that follow it
    int has_lo, has_hi;
    int lo_rows, hi_rows, k_last;
};

// ---------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------
// What the launchers need beyond Corr1dArgs when the storage type differs from
// the arithmetic type (integer box filter): taps are 0/1 floats, the epilogue
// divides by `size`.
struct BoxInfo {
    bool on = false;
    float qinv = 0.f, qc0 = 0.f;
};

template <typename T, typename TS>
static const char *kernel_name(bool strided, bool ldg = false)
{
    if (ldg) {   // the tiled kernels staged with plain loads (rows / bases not 16-byte multiples)
        if (sizeof(TS) != sizeof(T))
            return strided ? (sizeof(TS) == 2 ? "uniform1d_ldg_strided_u16" : "uniform1d_ldg_strided_u8")
                           : (sizeof(TS) == 2 ? "uniform1d_ldg_contig_u16" : "uniform1d_ldg_contig_u8");
        return strided ? (sizeof(T) == 4 ? "corr1d_ldg_strided_f32" : "corr1d_ldg_strided_f64")
                       : (sizeof(T) == 4 ? "corr1d_ldg_contig_f32" : "corr1d_ldg_contig_f64");
    }
    if (sizeof(TS) != sizeof(T))
        return strided ? (sizeof(TS) == 2 ? "uniform1d_tma_strided_u16" : "uniform1d_tma_strided_u8")
                       : (sizeof(TS) == 2 ? "uniform1d_tma_contig_u16" : "uniform1d_tma_contig_u8");
    return strided ? (sizeof(T) == 4 ? "corr1d_tma_strided_f32" : "corr1d_tma_strided_f64")
                   : (sizeof(T) == 4 ? "corr1d_tma_contig_f32" : "corr1d_tma_contig_f64");
}

// ---------------------------------------------------------------------------
// entry points of the kernel-family units (dtype = B200_F32 / F64 / U16 / U8 of
// the ARRAY; the integer dtypes are the exact box filter or min / max with
// float arithmetic).  -1 = geometry not eligible.
// ---------------------------------------------------------------------------
// corr1d_strided.cu
int tma_strided_corr(const Corr1dArgs &a, const BoxInfo &box);
int tma_strided_minmax(const Corr1dArgs &a, const BoxInfo &box, bool is_max);
int tma_peer_box(const Corr1dArgs &a, const BoxInfo &box);
// corr1d_contig.cu
int tma_contig_corr(const Corr1dArgs &a);
int tma_contig_box(const Corr1dArgs &a, const BoxInfo &box);
int tma_contig_minmax(const Corr1dArgs &a, bool is_max);

}  // namespace b200
