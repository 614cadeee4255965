/* A plain-C consumer of the C ABI -- TEST INFRASTRUCTURE.
 *
 * Proves, without Python and without a GPU, that include/b200_ndfilters.h is a C header (C99, no C++ or torch types
 * in any signature), that a C program links against libb200ndfilters.so, and that argument validation answers with
 * the documented status codes before any CUDA call is made (the codes map onto scipy's exception classes,
 * scipy/ndimage/_filters.py:600-608,1533-1540).  Built and run by tests/test_host_logic.py. */
#include <stdio.h>
#include <string.h>

#include "b200_ndfilters.h"

static int failures = 0;

static void expect(const char *what, int got, int want, const char *needle)
{
    const char *msg = b200_last_error();
    if (got != want || (needle && (!msg || !strstr(msg, needle)))) {
        printf("FAIL %s: status %d (want %d), message \"%s\" (want \"%s\")\n", what, got, want, msg ? msg : "(null)",
               needle ? needle : "");
        ++failures;
    }
}

int main(void)
{
    const int64_t shape[2] = {4, 4};
    const int64_t empty[2] = {0, 4};
    const int64_t wshape[2] = {3, 3};
    const int64_t origins[2] = {0, 0};
    const int64_t bad_origins[2] = {2, 0};
    const double w[9] = {1, 1, 1, 1, 1, 1, 1, 1, 1};
    void *in = (void *)256, *out = (void *)512;   /* never dereferenced: every call below fails validation */

    if (!b200_version() || !strstr(b200_version(), "sm_100a")) {
        printf("FAIL version string: %s\n", b200_version() ? b200_version() : "(null)");
        ++failures;
    }
    expect("origin outside the taps", b200_correlate1d(in, out, B200_F32, 2, shape, 0, w, 3, B200_MODE_REFLECT, 0.0, 2,
                                                        0, 0, B200_EPI_STORE, 0u, NULL), B200_EORIGIN, "origin");
    expect("unknown mode", b200_correlate1d(in, out, B200_F32, 2, shape, 0, w, 3, 7, 0.0, 0, 0, 0, B200_EPI_STORE, 0u,
                                             NULL), B200_EMODE, "boundary mode not supported");
    expect("unknown dtype", b200_correlate1d(in, out, 42, 2, shape, 0, w, 3, B200_MODE_REFLECT, 0.0, 0, 0, 0,
                                              B200_EPI_STORE, 0u, NULL), B200_EDTYPE, "data type not supported");
    expect("no weights", b200_correlate1d(in, out, B200_F32, 2, shape, 0, w, 0, B200_MODE_REFLECT, 0.0, 0, 0, 0,
                                           B200_EPI_STORE, 0u, NULL), B200_EINVAL, "no filter weights given");
    expect("in == out", b200_correlate1d(in, in, B200_F32, 2, shape, 0, w, 3, B200_MODE_REFLECT, 0.0, 0, 0, 0,
                                          B200_EPI_STORE, 0u, NULL), B200_EINVAL, "alias");
    expect("box size 0", b200_uniform_filter1d(in, out, B200_U16, 2, shape, 0, 0, B200_MODE_REFLECT, 0.0, 0, 0, 0, 0u,
                                                NULL), B200_EINVAL, "incorrect filter size");
    expect("box origin", b200_uniform_filter1d(in, out, B200_U16, 2, shape, 0, 3, B200_MODE_REFLECT, 0.0, 2, 0, 0, 0u,
                                                NULL), B200_EORIGIN, "invalid origin");
    expect("N-D origin", b200_correlate_nd(in, out, B200_F32, 2, shape, w, wshape, bad_origins, B200_MODE_WRAP, 0.0, 0,
                                            0, 0u, NULL), B200_EORIGIN, "origin");
    /* an empty array is a no-op, whatever the pointers */
    expect("empty array", b200_correlate1d(NULL, NULL, B200_F32, 2, empty, 0, w, 3, B200_MODE_REFLECT, 0.0, 0, 0, 0,
                                            B200_EPI_STORE, 0u, NULL), B200_OK, NULL);
    expect("empty array, N-D", b200_correlate_nd(NULL, NULL, B200_F32, 2, empty, w, wshape, origins, B200_MODE_WRAP,
                                                  0.0, 0, 0, 0u, NULL), B200_OK, NULL);
    /* host-only eligibility queries */
    if (b200_correlate1d_pair_eligible(B200_F64, 2, shape, 3, 0, 3, 0) != 0) {
        printf("FAIL float64 has no fused pair kernel\n");
        ++failures;
    }
    printf(failures ? "abi_check: %d failure(s)\n" : "abi_check: ok\n", failures);
    return failures ? 1 : 0;
}
