// CPU-only harness for the host side of lkb_expm.cu: the restated stdlib `expm` (Pade 10 + scaling and squaring) and the
// Fortran E16.9 formatter.  Includes the translation unit to reach its internal helpers; no CUDA call is made.
#include "lkb_expm.cu"
#include <cstdio>
#include <random>
int main() {
    // (1) the reference's own dense test (test/TestExpmlib.fypp test_dense_expm_*): A(i, i+1) = m, n = 5, m = 6:
    //     E(i, i+j) = m^j / j!
    {
        const int n = 5; const double m = 6.0;
        std::vector<cd> A((size_t)n * n, cd(0)), E;
        for (int i = 0; i < n - 1; ++i) A[i + (size_t)n * (i + 1)] = m;
        if (!dense_expm(n, A, E)) { printf("expm failed\n"); return 1; }
        double worst = 0;
        for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) {
            double ref = 0; if (j >= i) { ref = 1; for (int q = 1; q <= j - i; ++q) ref *= m / q; }
            worst = std::max(worst, std::abs(E[i + (size_t)n * j] - ref));
        }
        printf("nilpotent %.3e\n", worst);
    }
    // (2) exp(A) exp(-A) = I and exp of a skew-Hermitian matrix is unitary, n = 40, ||A|| ~ 10 (several squarings)
    {
        const int n = 40; std::mt19937 g(3); std::normal_distribution<double> nd;
        std::vector<cd> A((size_t)n * n), Am, S((size_t)n * n), E1, E2, P;
        for (auto& v : A) v = cd(nd(g), nd(g));
        Am = A; for (auto& v : Am) v = -v;
        if (!dense_expm(n, A, E1) || !dense_expm(n, Am, E2)) { printf("expm failed\n"); return 1; }
        matmul(n, E1, E2, P);
        double worst = 0, scale = 0;
        for (auto& v : E1) scale = std::max(scale, std::abs(v));
        for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) worst = std::max(worst, std::abs(P[i + (size_t)n * j] - (i == j ? 1.0 : 0.0)));
        printf("inverse %.3e scale %.3e\n", worst, scale);
        for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) S[i + (size_t)n * j] = A[i + (size_t)n * j] - std::conj(A[j + (size_t)n * i]);
        if (!dense_expm(n, S, E1)) { printf("expm failed\n"); return 1; }
        worst = 0;
        for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) {
            cd s = 0; for (int l = 0; l < n; ++l) s += std::conj(E1[l + (size_t)n * i]) * E1[l + (size_t)n * j];
            worst = std::max(worst, std::abs(s - (i == j ? 1.0 : 0.0)));
        }
        printf("unitary %.3e\n", worst);
    }
    // (3) E16.9 edit descriptor
    const double vals[] = {0.0, 1.0, -1.0, 0.1, 123456.789, -9.99999999999e-5, 1e100, 3.0e-310, 0.9999999996};
    for (double v : vals) printf("fmt [%s]\n", fortran_e16_9(v).c_str());
    // (4) matrix read from stdin (n, then n*n re/im pairs, column-major) -> exp printed, for comparison with scipy
    {
        int n = 0; if (scanf("%d", &n) != 1) return 0;
        std::vector<cd> A((size_t)n * n), E;
        for (auto& v : A) { double re, im; if (scanf("%lf %lf", &re, &im) != 2) return 1; v = cd(re, im); }
        if (!dense_expm(n, A, E)) { printf("expm failed\n"); return 1; }
        printf("matrix %d\n", n);
        for (auto& v : E) printf("%.17g %.17g\n", v.real(), v.imag());
    }
    return 0;
}
