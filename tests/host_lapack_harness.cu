// CPU-only harness for the host LAPACK layer of lkb_eig.cu (geev / gees + trsen / syev|heev / gesdd in the precision of the
// kind): includes the translation unit to reach its internal helpers and links liblkb.so for the rest.  No CUDA call is made,
// so it runs without a GPU (tests/test_host_lapack.py).  Prints one line per check: "<kind> <what> residual <r> ...".
#include "lkb_eig.cu"
#include <random>
#include <cstdio>
template <typename R> int run(bool cplx, const char* tag) {
    const int k = 24;
    std::mt19937 g(7); std::normal_distribution<double> nd;
    std::vector<cd> A((size_t)k * k);
    for (auto& v : A) v = cd((R)nd(g), cplx ? (R)nd(g) : 0.0);
    // eig: A v = lambda v
    std::vector<cd> vals, vecs;
    if (host_eig_t<R>(cplx, k, A, k, vals, vecs)) { printf("eig failed: %s\n", lkb_last_error()); return 1; }
    double worst = 0;
    for (int i = 0; i < k; ++i) {
        std::vector<cd> v(k);
        if (cplx || vals[i].imag() == 0) for (int r = 0; r < k; ++r) v[r] = vecs[r + (size_t)k * i];
        else if (vals[i].imag() > 0) for (int r = 0; r < k; ++r) v[r] = cd(vecs[r + (size_t)k * i].real(), vecs[r + (size_t)k * (i + 1)].real());
        else for (int r = 0; r < k; ++r) v[r] = cd(vecs[r + (size_t)k * (i - 1)].real(), -vecs[r + (size_t)k * i].real());
        for (int r = 0; r < k; ++r) { cd s = 0; for (int c = 0; c < k; ++c) s += A[r + (size_t)k * c] * v[c]; worst = std::max(worst, std::abs(s - vals[i] * v[r])); }
    }
    printf("%s eig residual %.2e\n", tag, worst);
    // schur select: A Z = Z T, Z unitary, nkeep about k/2, leading nkeep eigenvalues are the larger ones
    std::vector<cd> T, Z; int32_t nk = 0;
    if (host_schur_select_t<R>(cplx, k, A, T, Z, &nk)) { printf("schur failed: %s\n", lkb_last_error()); return 1; }
    worst = 0;
    for (int r = 0; r < k; ++r) for (int c = 0; c < k; ++c) { cd s1 = 0, s2 = 0; for (int l = 0; l < k; ++l) { s1 += A[r + (size_t)k * l] * Z[l + (size_t)k * c]; s2 += Z[r + (size_t)k * l] * T[l + (size_t)k * c]; } worst = std::max(worst, std::abs(s1 - s2)); }
    printf("%s schur residual %.2e nkeep %d\n", tag, worst, nk);
    // eigh of A + A^H
    std::vector<cd> S((size_t)k * k), vk; std::vector<double> ev(k);
    for (int r = 0; r < k; ++r) for (int c = 0; c < k; ++c) S[r + (size_t)k * c] = A[r + (size_t)k * c] + std::conj(A[c + (size_t)k * r]);
    if (host_eigh_t<R>(cplx, k, S, ev.data(), vk)) { printf("eigh failed\n"); return 1; }
    worst = 0;
    for (int i = 0; i < k; ++i) for (int r = 0; r < k; ++r) { cd s = 0; for (int c = 0; c < k; ++c) s += S[r + (size_t)k * c] * vk[c + (size_t)k * i]; worst = std::max(worst, std::abs(s - ev[i] * vk[r + (size_t)k * i])); }
    printf("%s eigh residual %.2e ascending %d\n", tag, worst, (int)std::is_sorted(ev.begin(), ev.end()));
    // svd: A V = U S
    std::vector<cd> uk, vv; std::vector<double> sv(k);
    if (host_svd_t<R>(cplx, k, A, sv.data(), uk, vv)) { printf("svd failed\n"); return 1; }
    worst = 0;
    for (int i = 0; i < k; ++i) for (int r = 0; r < k; ++r) { cd s = 0; for (int c = 0; c < k; ++c) s += A[r + (size_t)k * c] * vv[c + (size_t)k * i]; worst = std::max(worst, std::abs(s - sv[i] * uk[r + (size_t)k * i])); }
    printf("%s svd residual %.2e descending %d\n", tag, worst, (int)std::is_sorted(sv.rbegin(), sv.rend()));
    return 0;
}
int main(int argc, char** argv) {
    if (lapack_open(argv[1], "scipy_", "_")) { printf("open failed: %s\n", lkb_last_error()); return 1; }
    {   // sort_index(reverse=.true.): non-increasing, ties keep their original order
        const std::vector<double> key = {1.0, 3.0, 3.0, 2.0, 0.0, 0.0, 3.0};
        printf("sortidx");
        for (int i : sort_index_reverse(key)) printf(" %d", i);
        printf("\n");
    }
    return run<double>(false, "d") | run<double>(true, "z") | run<float>(false, "s") | run<float>(true, "c");
}
