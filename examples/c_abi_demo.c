/* c_abi_demo.c -- plain-C client of liblkb.so (no Python, no torch): what a Fortran/C host links against.
 *
 *   gcc -std=c99 -Iinclude examples/c_abi_demo.c -Llightkrylov_b200/csrc -llkb -Wl,-rpath,$PWD/lightkrylov_b200/csrc -lm -o c_abi_demo
 *
 * Runs arnoldi(kdim = 64) on the 5-point Poisson operator (config C2's operator on a 1024^2 grid), then
 * checks ||V^H V - I||_max with the device Gram matrix and the Arnoldi relation for the last column. */
#define _POSIX_C_SOURCE 200809L
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <time.h>
#include "lkb.h"

#define CHECK(call) do { int rc_ = (call); if (rc_ != 0) { fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, lkb_last_error()); return 2; } } while (0)

int main(int argc, char** argv) {
    const int64_t nx = argc > 1 ? atoll(argv[1]) : 1024, ny = nx, n = nx * ny;
    const int kdim = 64;
    const double coef[5] = {4.0, -1.0, -1.0, -1.0, -1.0};
    lkb_ctx_t ctx; lkb_op_t A; lkb_basis_t X; lkb_vec_t x0, y, r;
    CHECK(lkb_init(0, &ctx));
    CHECK(lkb_op_stencil5_create(ctx, LKB_D, nx, ny, coef, 0, ny, &A));
    CHECK(lkb_basis_create(ctx, LKB_D, n, n, 0, kdim + 1, &X));
    CHECK(lkb_basis_col(X, 0, &x0));
    CHECK(lkb_vec_fill_random(x0, LKB_DIST_UNIFORM, 42));
    double nrm; CHECK(lkb_vec_norm(x0, &nrm));
    double inv = 1.0 / nrm; CHECK(lkb_vec_scal(x0, &inv));

    double* H = (double*)calloc((size_t)(kdim + 1) * kdim, sizeof(double));
    int32_t info = -1;
    CHECK(lkb_arnoldi(A, X, H, kdim + 1, &info, 0, 0, -1.0, 0, 1));     /* warm-up: captures the step graph */
    struct timespec t0, t1; clock_gettime(CLOCK_MONOTONIC, &t0);
    CHECK(lkb_arnoldi(A, X, H, kdim + 1, &info, 0, 0, -1.0, 0, 1));
    clock_gettime(CLOCK_MONOTONIC, &t1);
    const double dt = (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);

    double* G = (double*)malloc(sizeof(double) * (kdim + 1) * (kdim + 1));
    CHECK(lkb_basis_innerprod(X, kdim + 1, X, 0, kdim + 1, G, kdim + 1));
    double orth = 0.0;
    for (int j = 0; j <= kdim; ++j) for (int i = 0; i <= kdim; ++i) {
        const double e = fabs(G[i + (kdim + 1) * j] - (i == j ? 1.0 : 0.0));
        if (e > orth) orth = e;
    }
    /* A v_k = V_{k+1} H(:, k) for k = kdim */
    lkb_vec_t vk;
    CHECK(lkb_vec_create(ctx, LKB_D, n, n, 0, &y)); CHECK(lkb_vec_create(ctx, LKB_D, n, n, 0, &r));
    CHECK(lkb_basis_col(X, kdim - 1, &vk));
    CHECK(lkb_op_matvec(A, vk, y));
    CHECK(lkb_basis_lincomb(X, kdim + 1, H + (size_t)(kdim + 1) * (kdim - 1), r));
    const double one = 1.0, mone = -1.0; double res;
    CHECK(lkb_vec_axpby(&mone, y, &one, r)); CHECK(lkb_vec_norm(r, &res));
    int64_t nmv = 0; CHECK(lkb_op_counters(A, &nmv, NULL));
    printf("c_abi_demo: n=%lld kdim=%d info=%d  %.1f steps/s  orth=%.2e  relation=%.2e  matvecs=%lld\n",
           (long long)n, kdim, (int)info, kdim / dt, orth, res, (long long)nmv);
    const int ok = info == 0 && orth < 1e-12 && res < 1e-11;
    lkb_vec_destroy(vk); lkb_vec_destroy(y); lkb_vec_destroy(r); lkb_vec_destroy(x0);
    lkb_basis_destroy(X); lkb_op_destroy(A); lkb_finalize(ctx);
    free(H); free(G);
    return ok ? 0 : 1;
}
