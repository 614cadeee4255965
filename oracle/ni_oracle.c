/*
 * oracle/ni_oracle.c -- TEST INFRASTRUCTURE ONLY (never imported by the product).
 *
 * A plain-C, single-threaded CPU restatement of the three native loops that
 * dask-image's ndfilters hot path ends in.  dask-image itself is pure Python;
 * per chunk it calls scipy.ndimage (reference call sites:
 * dask_image/dispatch/_dispatch_ndfilters.py:49,65,129,145,161,273), whose
 * Python drivers end in three compiled entry points of the third-party
 * dependency scipy (pyproject.toml:29 "scipy >=1.7.0"; 1.18.1 in this image):
 *
 *   _nd_image.correlate1d       (scipy/ndimage/_filters.py:610)
 *   _nd_image.uniform_filter1d  (scipy/ndimage/_filters.py:1542)
 *   _nd_image.correlate         (scipy/ndimage/_filters.py:1305)
 *
 * scipy's C sources are NOT present in this image (only the built extension),
 * so what follows restates the *published algorithm* (line buffers widened to
 * double, boundary extension by index map, accumulation order, plain C cast to
 * the array dtype) and is pinned bit-for-bit against the installed scipy in
 * tests/test_oracle.py and against the fixtures under tests/golden/.
 *
 * Parity status: PINNED (scipy 1.18.1 runs in this image and on the GPU box).
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -shared -fPIC).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* dtype codes shared with include/b200_ndfilters.h */
enum { DT_U8 = 0, DT_I8, DT_U16, DT_I16, DT_U32, DT_I32, DT_U64, DT_I64, DT_F32, DT_F64 };
/* scipy's extension-mode codes (scipy/ndimage/_ni_support.py:37-49) */
enum { M_NEAREST = 0, M_WRAP = 1, M_REFLECT = 2, M_MIRROR = 3, M_CONSTANT = 4 };

static size_t dt_size(int dt)
{
    switch (dt) {
    case DT_U8: case DT_I8: return 1;
    case DT_U16: case DT_I16: return 2;
    case DT_U32: case DT_I32: case DT_F32: return 4;
    default: return 8;
    }
}

static double load_as_double(const char *p, int dt)
{
    switch (dt) {
    case DT_U8:  return (double)*(const uint8_t *)p;
    case DT_I8:  return (double)*(const int8_t *)p;
    case DT_U16: return (double)*(const uint16_t *)p;
    case DT_I16: return (double)*(const int16_t *)p;
    case DT_U32: return (double)*(const uint32_t *)p;
    case DT_I32: return (double)*(const int32_t *)p;
    case DT_U64: return (double)*(const uint64_t *)p;
    case DT_I64: return (double)*(const int64_t *)p;
    case DT_F32: return (double)*(const float *)p;
    default:     return *(const double *)p;
    }
}

/* Plain C casts, exactly as the reference's compiled code stores a double line
 * into the array dtype (truncation toward zero for integers; out-of-range /
 * negative-to-unsigned values take whatever x86-64 gcc's cvttsd2si sequence
 * gives, which is what SURVEY.md Appendix A4 probed: u16 -1.5 -> 65535). */
static void store_from_double(char *p, int dt, double v)
{
    switch (dt) {
    case DT_U8:  *(uint8_t *)p = (uint8_t)v; break;
    case DT_I8:  *(int8_t *)p = (int8_t)v; break;
    case DT_U16: *(uint16_t *)p = (uint16_t)v; break;
    case DT_I16: *(int16_t *)p = (int16_t)v; break;
    case DT_U32: *(uint32_t *)p = (uint32_t)v; break;
    case DT_I32: *(int32_t *)p = (int32_t)v; break;
    case DT_U64: *(uint64_t *)p = (uint64_t)v; break;
    case DT_I64: *(int64_t *)p = (int64_t)v; break;
    case DT_F32: *(float *)p = (float)v; break;
    default:     *(double *)p = v; break;
    }
}

/* Index map of SURVEY.md Appendix A2: position i (any integer) on an axis of
 * length n -> source index in [0,n), or -1 for "use cval". */
static int64_t ext_index(int64_t i, int64_t n, int mode)
{
    if (i >= 0 && i < n) return i;
    switch (mode) {
    case M_NEAREST:
        return i < 0 ? 0 : n - 1;
    case M_WRAP: {
        int64_t t = i % n;
        return t < 0 ? t + n : t;
    }
    case M_REFLECT: {
        int64_t p = 2 * n;
        int64_t t = i % p;
        if (t < 0) t += p;
        return t < n ? t : p - 1 - t;
    }
    case M_MIRROR: {
        if (n == 1) return 0;
        int64_t p = 2 * n - 2;
        int64_t t = i % p;
        if (t < 0) t += p;
        return t < n ? t : p - t;
    }
    default:
        return -1;
    }
}

/* Gather one line of the array (along `axis`) into a double buffer extended by
 * `lo` samples before and `hi` after, per `mode`. */
static void fetch_line(const char *base, int dt, int64_t n, int64_t stride_bytes,
                       int64_t lo, int64_t hi, int mode, double cval, double *buf)
{
    for (int64_t k = -lo; k < n + hi; ++k) {
        int64_t s = ext_index(k, n, mode);
        buf[k + lo] = (s < 0) ? cval : load_as_double(base + s * stride_bytes, dt);
    }
}

/* Iterate over all lines of a C-contiguous array along `axis`:
 * line (o, c) starts at element o*n*inner + c, stride inner. */
static void line_geometry(int ndim, const int64_t *shape, int axis,
                          int64_t *outer, int64_t *n, int64_t *inner)
{
    int64_t o = 1, in = 1;
    for (int d = 0; d < axis; ++d) o *= shape[d];
    for (int d = axis + 1; d < ndim; ++d) in *= shape[d];
    *outer = o; *n = shape[axis]; *inner = in;
}

/*
 * correlate1d: out[i] = sum_j w[j] * ext(x)[i + j - nw/2 - origin]
 * (scipy/ndimage/_filters.py:556-612 is the driver).  Accumulation order of
 * the compiled loop (probe-verified): odd-length
 * weights that are symmetric / antisymmetric within DBL_EPSILON take a folded
 * branch starting from the centre tap then the outermost pair inward; anything
 * else starts from the LAST tap and then runs first..last-1.
 */
int oracle_correlate1d(const void *in, void *out, int dtype, int ndim,
                       const int64_t *shape, int axis, const double *w,
                       int64_t nw, int mode, double cval, int64_t origin)
{
    if (nw < 1 || axis < 0 || axis >= ndim) return 1;
    int64_t size1 = nw / 2, size2 = nw - size1 - 1;
    if (origin < -size1 || origin > (nw - 1) / 2) return 2;
    int64_t outer, n, inner;
    line_geometry(ndim, shape, axis, &outer, &n, &inner);
    if (outer * n * inner == 0) return 0;

    int symmetric = 0;
    if (nw & 1) {
        symmetric = 1;
        for (int64_t k = 1; k <= size1; ++k)
            if (fabs(w[size1 + k] - w[size1 - k]) > DBL_EPSILON) { symmetric = 0; break; }
        if (!symmetric) {
            symmetric = -1;
            for (int64_t k = 1; k <= size1; ++k)
                if (fabs(w[size1 + k] + w[size1 - k]) > DBL_EPSILON) { symmetric = 0; break; }
        }
    }
    int64_t lo = size1 + origin, hi = size2 - origin;
    size_t es = dt_size(dtype);
    double *buf = (double *)malloc(sizeof(double) * (size_t)(n + lo + hi));
    if (!buf) return 3;
    const double *fw = w + size1;
    for (int64_t o = 0; o < outer; ++o) {
        for (int64_t c = 0; c < inner; ++c) {
            size_t off = ((size_t)o * n * inner + c) * es;
            fetch_line((const char *)in + off, dtype, n, inner * es, lo, hi, mode, cval, buf);
            char *po = (char *)out + off;
            for (int64_t l = 0; l < n; ++l) {
                const double *x = buf + l + size1; /* centre sample of the window */
                double acc;
                if (symmetric > 0) {
                    acc = x[0] * fw[0];
                    for (int64_t j = -size1; j < 0; ++j) acc += (x[j] + x[-j]) * fw[j];
                } else if (symmetric < 0) {
                    acc = x[0] * fw[0];
                    for (int64_t j = -size1; j < 0; ++j) acc += (x[j] - x[-j]) * fw[j];
                } else {
                    acc = x[size2] * fw[size2];
                    for (int64_t j = -size1; j < size2; ++j) acc += x[j] * fw[j];
                }
                store_from_double(po + l * inner * es, dtype, acc);
            }
        }
    }
    free(buf);
    return 0;
}

/*
 * uniform_filter1d (scipy/ndimage/_filters.py:1502-1549): double running sum
 * per line, out = sum / size (the order
 * "tmp += new - old; out = tmp / size" was re-probed bit-exactly here).
 */
int oracle_uniform_filter1d(const void *in, void *out, int dtype, int ndim,
                            const int64_t *shape, int axis, int64_t size,
                            int mode, double cval, int64_t origin)
{
    if (size < 1 || axis < 0 || axis >= ndim) return 1;
    int64_t size1 = size / 2, size2 = size - size1 - 1;
    if (size1 + origin < 0 || size1 + origin >= size) return 2;
    int64_t outer, n, inner;
    line_geometry(ndim, shape, axis, &outer, &n, &inner);
    if (outer * n * inner == 0) return 0;
    int64_t lo = size1 + origin, hi = size2 - origin;
    size_t es = dt_size(dtype);
    double *buf = (double *)malloc(sizeof(double) * (size_t)(n + lo + hi));
    if (!buf) return 3;
    for (int64_t o = 0; o < outer; ++o) {
        for (int64_t c = 0; c < inner; ++c) {
            size_t off = ((size_t)o * n * inner + c) * es;
            fetch_line((const char *)in + off, dtype, n, inner * es, lo, hi, mode, cval, buf);
            char *po = (char *)out + off;
            double tmp = 0.0;
            for (int64_t k = 0; k < size; ++k) tmp += buf[k];
            store_from_double(po, dtype, tmp / (double)size);
            for (int64_t l = 1; l < n; ++l) {
                tmp += buf[l + size - 1] - buf[l - 1];
                store_from_double(po + l * inner * es, dtype, tmp / (double)size);
            }
        }
    }
    free(buf);
    return 0;
}

/*
 * N-D correlate (driver: scipy/ndimage/_filters.py:1253-1309).  One pass over
 * the output voxels; tmp = 0; for each weight with |w| > DBL_EPSILON, in C
 * order over the weights array: tmp += w * (double)x[neighbour or cval]
 * (SURVEY.md section 3.4).  Neighbour position along axis a for weight index
 * k_a is  i_a + k_a - wshape_a/2 - origin_a, extended per `mode`.
 */
#define ORACLE_MAXDIM 8
int oracle_correlate_nd(const void *in, void *out, int dtype, int ndim,
                        const int64_t *shape, const double *w,
                        const int64_t *wshape, const int64_t *origins,
                        int mode, double cval)
{
    if (ndim < 1 || ndim > ORACLE_MAXDIM) return 1;
    int64_t total = 1, wtotal = 1;
    int64_t strides[ORACLE_MAXDIM];
    for (int d = 0; d < ndim; ++d) {
        total *= shape[d]; wtotal *= wshape[d];
        int64_t s1 = wshape[d] / 2;
        if (wshape[d] < 1) return 1;
        if (origins[d] < -s1 || origins[d] > (wshape[d] - 1) / 2) return 2;
    }
    if (total == 0) return 0;
    strides[ndim - 1] = 1;
    for (int d = ndim - 2; d >= 0; --d) strides[d] = strides[d + 1] * shape[d + 1];
    /* compact list of the non-zero taps, C order */
    int64_t nnz = 0;
    double *tw = (double *)malloc(sizeof(double) * (size_t)wtotal);
    int64_t *tk = (int64_t *)malloc(sizeof(int64_t) * (size_t)wtotal * ndim);
    if (!tw || !tk) { free(tw); free(tk); return 3; }
    for (int64_t f = 0; f < wtotal; ++f) {
        if (fabs(w[f]) > DBL_EPSILON) {
            int64_t r = f;
            for (int d = ndim - 1; d >= 0; --d) {
                tk[nnz * ndim + d] = r % wshape[d] - wshape[d] / 2 - origins[d];
                r /= wshape[d];
            }
            tw[nnz++] = w[f];
        }
    }
    size_t es = dt_size(dtype);
    int64_t idx[ORACLE_MAXDIM];
    memset(idx, 0, sizeof(idx));
    for (int64_t p = 0; p < total; ++p) {
        double tmp = 0.0;
        for (int64_t t = 0; t < nnz; ++t) {
            int64_t off = 0;
            int outside = 0;
            for (int d = 0; d < ndim; ++d) {
                int64_t s = ext_index(idx[d] + tk[t * ndim + d], shape[d], mode);
                if (s < 0) { outside = 1; break; }
                off += s * strides[d];
            }
            double v = outside ? cval : load_as_double((const char *)in + (size_t)off * es, dtype);
            tmp += tw[t] * v;
        }
        store_from_double((char *)out + (size_t)p * es, dtype, tmp);
        for (int d = ndim - 1; d >= 0; --d) {
            if (++idx[d] < shape[d]) break;
            idx[d] = 0;
        }
    }
    free(tw); free(tk);
    return 0;
}
