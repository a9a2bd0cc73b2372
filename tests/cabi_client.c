/* A plain C99 client of the C-ABI (include/sdfb200.h): no Python, no torch, no CUDA headers.
 *
 * What a binder in any host language does: build a program image by hand (sphere(1), d3.py:92-96 -- opcode
 * numbers from csrc/sdf_opcodes.h), the float64 axes as np.arange produces them (core.py:110-112), then the
 * count pass, a page-locked result buffer and the real call.  Used by tests/test_host_cpu.py (no GPU: the call
 * must fail with SDFB200_E_NOGPU and say so) and tests/test_gpu_zzz_cabi_client.py (result == the Python path's).
 *
 *   cabi_client <n> <batch_size>      lattice of n^3 samples over [-1.25, 1.25)^3
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "sdfb200.h"
#include "sdf_opcodes.h"

static int fail(const char* what, int rc) {
    printf("FAIL %s rc=%d error=%s\n", what, rc, sdfb200_last_error());
    return 1;
}

int main(int argc, char** argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 64;
    const int batch_size = argc > 2 ? atoi(argv[2]) : 32;
    if (sdfb200_abi_version() != 2) return fail("abi_version", sdfb200_abi_version());

    /* malformed programs and NULL arguments are refused with a message, not a crash */
    sdfb200_prog* prog = NULL;
    const uint32_t bad_code[1] = { 0xffu };
    const double no_consts[1] = { 0.0 };
    int rc = sdfb200_program_create(bad_code, 1, no_consts, 0, &prog);
    if (rc != SDFB200_E_INVALID || strlen(sdfb200_last_error()) == 0) return fail("bad opcode accepted", rc);
    rc = sdfb200_program_create(NULL, 1, no_consts, 0, &prog);
    if (rc != SDFB200_E_INVALID) return fail("NULL code accepted", rc);

    /* sphere(1): one word, constants cx cy cz r */
    const uint32_t code[1] = { (uint32_t)OP_SPHERE | (0u << 8) | (0u << 16) };
    const double consts[4] = { 0.0, 0.0, 0.0, 1.0 };
    rc = sdfb200_program_create(code, 1, consts, 4, &prog);
    if (rc) return fail("program_create", rc);

    double* axis = (double*)malloc(sizeof(double) * (size_t)n);
    const double x0 = -1.25, step = 2.5 / n;
    for (int i = 0; i < n; ++i) axis[i] = x0 + i * step;            /* np.arange(x0, x1, step) */

    int64_t n_tri = -1, stats[6] = { 0 };
    rc = sdfb200_generate(prog, axis, n, axis, n, axis, n, batch_size, 1, SDFB200_PREC_F64, 0.0, 0, -1,
                          NULL, 0, &n_tri, stats, NULL, 0, NULL);   /* count pass */
    if (rc == SDFB200_E_NOGPU) {
        printf("NOGPU error=%s\n", sdfb200_last_error());
        sdfb200_program_free(prog);
        free(axis);
        return 0;
    }
    if (rc != 0 && rc != SDFB200_E_CAPACITY) return fail("generate (count)", rc);
    const int64_t need = n_tri;

    void* buf = NULL;
    rc = sdfb200_host_alloc((size_t)(need > 0 ? need : 1) * 36, &buf);
    if (rc) return fail("host_alloc", rc);
    if (need > 1) {                                                  /* too small: error code, never truncation */
        rc = sdfb200_generate(prog, axis, n, axis, n, axis, n, batch_size, 1, SDFB200_PREC_F64, 0.0, 0, -1,
                              (float*)buf, need - 1, &n_tri, stats, NULL, 0, NULL);
        if (rc != SDFB200_E_CAPACITY || n_tri != need) return fail("capacity overflow not reported", rc);
    }
    rc = sdfb200_generate(prog, axis, n, axis, n, axis, n, batch_size, 1, SDFB200_PREC_F64, 0.0, 0, -1,
                          (float*)buf, need, &n_tri, stats, NULL, 0, NULL);
    if (rc) return fail("generate", rc);

    /* order-independent digest of the soup: sums of the coordinates and of their squares, in double */
    const float* v = (const float*)buf;
    double s1 = 0.0, s2 = 0.0, rmin = 1e30, rmax = 0.0;
    for (int64_t i = 0; i < n_tri * 3; ++i) {
        const double x = v[3 * i], y = v[3 * i + 1], z = v[3 * i + 2], r2 = x * x + y * y + z * z;
        s1 += x + y + z; s2 += r2;
        if (r2 < rmin) rmin = r2;
        if (r2 > rmax) rmax = r2;
    }
    printf("OK n_tri=%lld skipped=%lld empty=%lld nonempty=%lld sum=%.17g sumsq=%.17g r2min=%.9g r2max=%.9g\n",
           (long long)n_tri, (long long)stats[0], (long long)stats[1], (long long)stats[2], s1, s2, rmin, rmax);
    sdfb200_host_free(buf);
    sdfb200_program_free(prog);
    free(axis);
    return 0;
}
