/* lk_oracle.c -- CPU oracle for the LightKrylov hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the checker, never the product: only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * (lightkrylov_b200/csrc) never links, imports or executes anything under oracle/.
 *
 * It is a plain-C restatement of the reference's per-vector algorithm (see
 * lko_body.inc for the file:line map).  The Fortran reference cannot be compiled in
 * this image (no Fortran compiler, fpm, fypp or stdlib), so there
 * is no oracle/_ref.
 *
 * PINNING.  The reference holds no golden vectors or fixtures for this path (its tests are
 * property / known-answer checks on unseeded random inputs) and it cannot be COMPILED here or
 * on the GPU box (both probed: no gfortran / flang / ifx / nvfortran / fpm / fypp).  Since
 * round 2 its source text is EXECUTED instead: oracle/f90run.py interprets the reference's
 * pre-expanded .f90 files (modules, submodules, type-bound procedures, generics) on the
 * reference's own TestUtils vector / operator types; tests/golden/ref_krylov.npz and
 * ref_solvers.npz hold what its arnoldi / lanczos / bidiagonalization / qr / Gram-Schmidt /
 * gmres / fgmres / cg / eigs / eighs / svds / kexpm code computed (generating script
 * tests/golden/make_ref_golden.py), and tests/test_ref_golden.py requires THIS code to
 * reproduce them in all four kinds (fp64 1e-12, fp32 5e-5; info, pivots, iteration counts
 * exact).  The interpreter itself is qualified by the reference's own 112 unit tests
 * passing under it (tests/test_reference_suite.py).  What that leaves open: it is an
 * interpreter, not a compiled build -- intrinsics, stdlib BLAS / LAPACK / expm are supplied
 * by numpy / scipy, so summation order inside a dot product is numpy's, not gfortran's.
 * Earlier pins stay in place: (1) the reference's own property / known-answer assertions
 * (test/TestKrylov.fypp:194-514, test/TestIterativeSolvers.fypp:59-725) re-run against this
 * code by tests/test_oracle_pins.py; (2) an independent evaluation of the Hessenberg entries
 * in 80-bit extended precision; (3) line-by-line citation of the reference in lko_body.inc;
 * (4) a second numpy restatement written from the Fortran sources
 * (tests/test_oracle_second_opinion.py) that must agree with this code entry by entry.
 *
 * Build:  make -C oracle      (gcc -O3 -march=native -fopenmp -shared)
 */
#include <complex.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ---- counter-based RNG shared with the device (lightkrylov_b200/csrc/lkb_rng.h) ----
 * h(seed, row, stream) = splitmix64 finaliser of a 64-bit counter; uniform in (0,1) with
 * 53 bits; normal by Box-Muller on streams (2c, 2c+1) for component c (0 = re, 1 = im). */
static inline uint64_t lko_mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
static inline double lko_u01(uint64_t seed, uint64_t row, uint64_t stream) {
    uint64_t h = lko_mix64(lko_mix64(seed) ^ (row * 4ULL + stream));
    return ((double)(h >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}
static inline double lko_rng_uniform(uint64_t seed, uint64_t row, int comp) {
    return lko_u01(seed, row, 2ULL * (uint64_t)comp);
}
static inline double lko_rng_normal(uint64_t seed, uint64_t row, int comp) {
    double u1 = lko_u01(seed, row, 2ULL * (uint64_t)comp);
    double u2 = lko_u01(seed, row, 2ULL * (uint64_t)comp + 1ULL);
    return sqrt(-2.0 * log(u1)) * cos(6.283185307179586476925286766559 * u2);
}

void lko_set_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n > 0 ? n : 1);
#else
    (void)n;
#endif
}
int lko_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* Constants.f90:16-48  atol = 10^-precision, rtol = sqrt(atol) */
#define T float
#define R float
#define SFX s
#define IS_CPLX 0
#define SFX_FABS fabsf
#define SFX_SQRT sqrtf
#define SFX_ATOL 1e-6f
#include "lko_body.inc"
#undef T
#undef R
#undef SFX
#undef IS_CPLX
#undef SFX_FABS
#undef SFX_SQRT
#undef SFX_ATOL

#define T double
#define R double
#define SFX d
#define IS_CPLX 0
#define SFX_FABS fabs
#define SFX_SQRT sqrt
#define SFX_ATOL 1e-15
#include "lko_body.inc"
#undef T
#undef R
#undef SFX
#undef IS_CPLX
#undef SFX_FABS
#undef SFX_SQRT
#undef SFX_ATOL

#define T float complex
#define R float
#define SFX c
#define IS_CPLX 1
#define SFX_CONJ conjf
#define SFX_CABS cabsf
#define SFX_CREAL crealf
#define SFX_CIMAG cimagf
#define SFX_SQRT sqrtf
#define SFX_ATOL 1e-6f
#include "lko_body.inc"
#undef T
#undef R
#undef SFX
#undef IS_CPLX
#undef SFX_CONJ
#undef SFX_CABS
#undef SFX_CREAL
#undef SFX_CIMAG
#undef SFX_SQRT
#undef SFX_ATOL

#define T double complex
#define R double
#define SFX z
#define IS_CPLX 1
#define SFX_CONJ conj
#define SFX_CABS cabs
#define SFX_CREAL creal
#define SFX_CIMAG cimag
#define SFX_SQRT sqrt
#define SFX_ATOL 1e-15
#include "lko_body.inc"
