/*
 * b200_ndfilters.h -- C ABI of the B200-native ndfilters correlation path.
 *
 * This is the drop-in boundary for the hot path of dask-image's ndfilters: the
 * per-chunk arithmetic that the reference reaches through scipy's three
 * compiled entry points
 *
 *   _nd_image.correlate1d(input, weights, axis, output, mode, cval, origin)
 *        scipy/ndimage/_filters.py:610   (reached from dask-image via
 *        dask_image/dispatch/_dispatch_ndfilters.py:129,145,161 ->
 *        gaussian_filter / gaussian_gradient_magnitude / gaussian_laplace)
 *   _nd_image.uniform_filter1d(input, size, axis, output, mode, cval, origin)
 *        scipy/ndimage/_filters.py:1542  (via _dispatch_ndfilters.py:273)
 *   _nd_image.correlate(input, weights, output, mode, cval, origins)
 *        scipy/ndimage/_filters.py:1305  (via _dispatch_ndfilters.py:49,65)
 *
 * plus the numpy ufunc epilogues scipy applies between passes
 * (scipy/ndimage/_filters.py:1025-1030 and :1184-1193).
 *
 * Conventions
 *  - plain pointers and sizes only; every array is C-contiguous, `shape` has
 *    `ndim` int64 entries, element type given by a b200_dtype code;
 *  - `in` / `out` are DEVICE pointers on the current CUDA device, `weights`,
 *    `shape`, `wshape`, `origins` are HOST pointers (copied before return);
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *    calls are asynchronous on that stream and re-entrant (no global mutable
 *    state besides a thread-local error string);
 *  - `in` and `out` must not overlap;
 *  - mode codes are scipy's (scipy/ndimage/_ni_support.py:37-49);
 *  - every function returns 0 on success or a B200_E* code; the message is
 *    available from b200_last_error() on the calling thread.
 *
 * Slab halos (multi-GPU axis-0 decomposition, replaces dask's map_overlap):
 * `in` / `out` point at the first of the `shape[0]` planes this call produces,
 * and `pad_lo` / `pad_hi` say how many REAL planes of the global volume are
 * resident in memory directly before `in` plane 0 / after plane n-1 along
 * axis 0 (own planes of the slab, or planes received from the neighbouring
 * rank).  Positions inside the resident range [-pad_lo, n + pad_hi) are read
 * as they are; positions outside it are folded by `mode` at the ends of the
 * resident range (on an end slab that end is the global boundary; elsewhere
 * the pads cover the filter's reach and it is never touched).  pad_lo =
 * pad_hi = 0 gives exactly scipy's behaviour on a whole array.  Pads are only
 * meaningful for the filtered axis 0 (1-D passes with axis == 0, and the N-D
 * correlate).
 */
#ifndef B200_NDFILTERS_H
#define B200_NDFILTERS_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200_MAXDIM 8

typedef enum {
    B200_U8 = 0, B200_I8 = 1, B200_U16 = 2, B200_I16 = 3, B200_U32 = 4,
    B200_I32 = 5, B200_U64 = 6, B200_I64 = 7, B200_F32 = 8, B200_F64 = 9
} b200_dtype;

typedef enum {
    B200_MODE_NEAREST = 0, B200_MODE_WRAP = 1, B200_MODE_REFLECT = 2,
    B200_MODE_MIRROR = 3, B200_MODE_CONSTANT = 4
} b200_mode;

/* status codes */
#define B200_OK            0
#define B200_EINVAL        1   /* bad argument (maps to RuntimeError)        */
#define B200_EORIGIN       2   /* origin outside the footprint (ValueError)  */
#define B200_EMODE         3   /* boundary mode not supported (RuntimeError) */
#define B200_EDTYPE        4   /* data type not supported (RuntimeError)     */
#define B200_ECUDA         5   /* CUDA runtime / driver error                */
#define B200_ENODEVICE     6   /* no sm_100 device / kernels unavailable     */

/* flags */
#define B200_FLAG_EXACT      1u  /* force the one-thread-per-output scipy-order fp64 kernels: bit-exact for
                                   * floats too; for 8/16-bit integers the tiled kernels are bit-exact as
                                   * well and this flag selects their cross-check */
#define B200_FLAG_NO_TMA     2u  /* never pick a TMA-tiled kernel (debug/tests) */
#define B200_FLAG_FORCE_FAST 4u  /* fail instead of falling back to generic     */

/* epilogue applied when a 1-D pass stores its result r (already rounded to
 * the array dtype, as scipy stores every pass) into out: */
typedef enum {
    B200_EPI_STORE = 0,        /* out  = r                          */
    B200_EPI_ACC = 1,          /* out += r            (laplace sum) */
    B200_EPI_SQ_STORE = 2,     /* out  = r*r          (ggm first)   */
    B200_EPI_SQ_ACC = 3,       /* out += r*r          (ggm middle)  */
    B200_EPI_SQ_ACC_SQRT = 4   /* out  = sqrt(out + r*r) (ggm last) */
} b200_epilogue;

/* replaces _nd_image.correlate1d (scipy/ndimage/_filters.py:610):
 *   out[i] = sum_j weights[j] * ext(in)[i + j - nw/2 - origin] along `axis` */
int b200_correlate1d(const void *in, void *out, int dtype, int ndim,
                     const int64_t *shape, int axis, const double *weights,
                     int64_t nw, int mode, double cval, int64_t origin,
                     int64_t pad_lo, int64_t pad_hi, int epilogue,
                     unsigned flags, void *stream);

/* The same pass along axis 0 of a slab whose neighbour planes are NOT resident
 * next to it: `halo_lo` points at the `n_lo` planes that precede plane 0 of the
 * slab in the global volume, `halo_hi` at the `n_hi` planes that follow its last
 * plane -- anywhere in device-accessible memory, typically a peer GPU's slab
 * mapped through CUDA IPC, read over NVLink by the kernel's own TMA loads
 * (multi-GPU slab engine, replaces dask's map_overlap halo assembly without an
 * exchange step).  n_lo / n_hi must equal the filter's reach below / above
 * (nw/2 + origin, nw - 1 - nw/2 - origin) or be 0 with a NULL pointer, which
 * makes that end of the slab a boundary of the volume (`mode` applies).
 * float32 / float64 only; returns B200_EINVAL ("not eligible") when the tile
 * geometry cannot place the neighbour rows -- callers then copy the planes next
 * to the slab and use b200_correlate1d with pad_lo / pad_hi. */
int b200_correlate1d_halo(const void *in, void *out, int dtype, int ndim,
                          const int64_t *shape, const double *weights,
                          int64_t nw, int mode, double cval, int64_t origin,
                          const void *halo_lo, int64_t n_lo,
                          const void *halo_hi, int64_t n_hi, int epilogue,
                          unsigned flags, void *stream);

/* Two consecutive b200_correlate1d passes in ONE launch: along axis ndim-2 with
 * taps `w1` (mode1, origin1), then along axis ndim-1 with taps `w2` (mode2,
 * origin2); `epilogue` applies to the store of the second pass.  This is the
 * tail of scipy's gaussian_filter loop (scipy/ndimage/_filters.py:852-858:
 * one correlate1d per axis, each pass stored in the array dtype): the
 * intermediate pass lives in shared memory, rounded to float32 exactly as the
 * stored pass would be, and both passes sum in the order of the one-pass
 * kernels, so the result is bit-identical to the two calls -- at 8 B/voxel of
 * HBM traffic instead of 16.  float32 only, ndim >= 2; rows that are not
 * 16-byte multiples are staged with plain loads instead of TMA.  Returns
 * B200_EINVAL ("not eligible") when no fused kernel exists for this pair of
 * filters; the _eligible query is host-only and depends on the dtype, the last
 * two extents and the filters alone (never on the leading extents, so every
 * slab of a volume takes the same path). */
int b200_correlate1d_pair(const void *in, void *out, int dtype, int ndim,
                          const int64_t *shape, const double *w1, int64_t nw1,
                          int mode1, int64_t origin1, const double *w2,
                          int64_t nw2, int mode2, int64_t origin2, double cval,
                          int epilogue, unsigned flags, void *stream);
int b200_correlate1d_pair_eligible(int dtype, int ndim, const int64_t *shape,
                                   int64_t nw1, int64_t origin1, int64_t nw2,
                                   int64_t origin2);

/* Let kernels on the current device read memory of `peer_device` (needed once
 * per device pair before b200_correlate1d_halo is given a peer GPU's planes;
 * cudaDeviceEnablePeerAccess, "already enabled" is success). */
int b200_enable_peer_access(int peer_device);

/* Sharing a slab with the neighbouring ranks (one process per GPU) so that their
 * axis-0 passes can read its edge planes in place (b200_correlate1d_halo):
 * b200_ipc_export fills the 64-byte CUDA IPC handle of the ALLOCATION that holds
 * `dev_ptr` and the byte offset of `dev_ptr` inside it (device buffers come from a
 * caching allocator: the pointer is rarely the allocation's base); the consumer
 * calls b200_ipc_open with the handle on ITS current device -- the mapping is then
 * valid for kernels of that device (peer access is enabled as needed) -- adds the
 * offset, and calls b200_ipc_close(base) when the slab is retired.  A process
 * cannot open its own handles.  Allocations made with CUDA's virtual memory API
 * (e.g. PYTORCH_CUDA_ALLOC_CONF=expandable_segments) cannot be exported:
 * B200_ECUDA, and the callers fall back to the NCCL exchange. */
#define B200_IPC_HANDLE_BYTES 64
int b200_ipc_export(const void *dev_ptr, void *handle_out, int64_t *offset_out);
int b200_ipc_open(const void *handle, int owner_device, void **base_out);
int b200_ipc_close(void *base);

/* 1 if b200_correlate1d_halo can take an axis-0 pass over `n` planes of `inner`
 * elements each with these taps, else 0 (host-only query, no CUDA call: every
 * rank of a slab decomposition evaluates it for every slab size so that all of
 * them choose the same path) */
int b200_correlate1d_halo_eligible(int dtype, int64_t n, int64_t inner,
                                   int64_t nw, int64_t origin);

/* replaces _nd_image.uniform_filter1d (scipy/ndimage/_filters.py:1542) */
int b200_uniform_filter1d(const void *in, void *out, int dtype, int ndim,
                          const int64_t *shape, int axis, int64_t size,
                          int mode, double cval, int64_t origin,
                          int64_t pad_lo, int64_t pad_hi, unsigned flags,
                          void *stream);

/* b200_uniform_filter1d along axis 0 of a slab with the neighbour planes read in
 * place (see b200_correlate1d_halo): float32 / float64, and the exact uint8 /
 * uint16 box filter.  The _eligible query is host-only. */
int b200_uniform_filter1d_halo(const void *in, void *out, int dtype, int ndim,
                               const int64_t *shape, int64_t size, int mode,
                               double cval, int64_t origin, const void *halo_lo,
                               int64_t n_lo, const void *halo_hi, int64_t n_hi,
                               unsigned flags, void *stream);
int b200_uniform_filter1d_halo_eligible(int dtype, int64_t n, int64_t inner,
                                        int64_t size, int64_t origin, int mode,
                                        double cval);

/* Two consecutive b200_uniform_filter1d passes in ONE launch: along axis ndim-2
 * (size1, mode1, origin1), then along axis ndim-1 (size2, mode2, origin2) -- the
 * tail of scipy's uniform_filter loop (scipy/ndimage/_filters.py:1613-1617: one
 * uniform_filter1d per axis, each pass stored in the array dtype).  uint8 /
 * uint16 only: the first pass's trunc(sum / size) is taken in shared memory
 * exactly as the stored pass would hold it, so the result is bit-identical to
 * the two calls at half their HBM traffic.  Needs equal odd sizes, origin 0,
 * rows that are 16-byte multiples and 16-byte aligned pointers; returns
 * B200_EINVAL ("not eligible") otherwise.  The _eligible query is host-only and
 * depends on the dtype, the last two extents and the filters alone. */
int b200_uniform_filter1d_pair(const void *in, void *out, int dtype, int ndim,
                               const int64_t *shape, int64_t size1, int mode1,
                               int64_t origin1, int64_t size2, int mode2,
                               int64_t origin2, double cval, unsigned flags,
                               void *stream);
int b200_uniform_filter1d_pair_eligible(int dtype, int ndim,
                                        const int64_t *shape, int64_t size1,
                                        int64_t origin1, int mode1,
                                        int64_t size2, int64_t origin2,
                                        int mode2, double cval);

/* replaces _nd_image.correlate (scipy/ndimage/_filters.py:1305); `weights` has
 * shape `wshape` (ndim entries, C order), taps with |w| <= DBL_EPSILON are
 * skipped as scipy does; convolve = flipped weights + adjusted origins,
 * prepared by the caller exactly as scipy/ndimage/_filters.py:1282-1287. */
int b200_correlate_nd(const void *in, void *out, int dtype, int ndim,
                      const int64_t *shape, const double *weights,
                      const int64_t *wshape, const int64_t *origins, int mode,
                      double cval, int64_t pad_lo, int64_t pad_hi,
                      unsigned flags, void *stream);

/* b200_correlate_nd over a slab of a 3-D float32 volume (planes along axis 0) whose
 * neighbour planes are read in place, as b200_correlate1d_halo does for the 1-D
 * passes: `halo_lo` points at the n_lo = wshape[0]/2 + origins[0] planes that
 * precede plane 0, `halo_hi` at the n_hi = wshape[0] - 1 - n_lo planes that follow
 * the last one (peer GPU memory mapped through b200_ipc_open), or NULL / 0 where
 * the slab ends at a boundary of the volume.  Returns B200_EINVAL ("not eligible")
 * when the tile geometry cannot place the neighbour planes; the _eligible query is
 * host-only (all ranks evaluate it for every slab size and choose alike). */
int b200_correlate_nd_halo(const void *in, void *out, int dtype, int ndim,
                           const int64_t *shape, const double *weights,
                           const int64_t *wshape, const int64_t *origins,
                           int mode, double cval, const void *halo_lo,
                           int64_t n_lo, const void *halo_hi, int64_t n_hi,
                           unsigned flags, void *stream);
int b200_correlate_nd_halo_eligible(int dtype, int ndim, const int64_t *shape,
                                    const int64_t *wshape,
                                    const int64_t *origins);

/* replaces the numpy ufuncs of scipy/ndimage/_filters.py:1025-1030,1184-1193:
 * acc (op)= src elementwise over n elements, in the array dtype. */
int b200_combine(void *acc, const void *src, int dtype, int64_t n, int op,
                 void *stream);

/* replaces _nd_image.min_or_max_filter1d (scipy/ndimage/_filters.py:1660-1666,
 * 1720-1726; reached via dask_image/dispatch/_dispatch_ndfilters.py:193,209):
 * minimum (is_max = 0) or maximum (is_max != 0) of the `size` window samples */
int b200_minmax_filter1d(const void *in, void *out, int dtype, int ndim,
                         const int64_t *shape, int axis, int64_t size,
                         int is_max, int mode, double cval, int64_t origin,
                         int64_t pad_lo, int64_t pad_hi, unsigned flags,
                         void *stream);

/* replaces _nd_image.min_or_max_filter (scipy/ndimage/_filters.py:1771) for a
 * boolean `footprint` of shape `fshape` (host, one byte per element, C order) */
int b200_minmax_filter_nd(const void *in, void *out, int dtype, int ndim,
                          const int64_t *shape, const uint8_t *footprint,
                          const int64_t *fshape, const int64_t *origins,
                          int is_max, int mode, double cval, int64_t pad_lo,
                          int64_t pad_hi, unsigned flags, void *stream);

/* frees the small device tables the library caches (compacted tap tables of the
 * generic and integer N-D kernels); optional, safe to call at any time */
int b200_shutdown(void);

/* name of the kernel family the last successful call on this thread chose
 * ("corr1d_tma_strided_f32", "corr1d_generic", ...); for tests and profiling */
const char *b200_last_kernel(void);
const char *b200_last_error(void);
const char *b200_version(void);
/* number of kernel launches issued by this thread since the last reset */
int64_t b200_launch_count(int reset);

#ifdef __cplusplus
}
#endif
#endif /* B200_NDFILTERS_H */
