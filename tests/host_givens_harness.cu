// CPU-only harness for the host Givens helpers of lkb_solvers.cu (apply_givens_rotation of the reference:
// submodule_utility_functions.fypp:169-204): `lartg` restates LAPACK 3.10 la_lartg and is compared with the provider's
// dlartg; givens_real = lasr('L','V','F') + lartg; givens_cplx = the reference's hand-rolled form.  No CUDA call is made.
#include "lkb_solvers.cu"
#include <dlfcn.h>
#include <cstdio>
#include <random>
int main(int argc, char** argv) {
    void* h = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL);
    if (!h) { printf("dlopen failed\n"); return 1; }
    typedef void (*lartg_t)(const double*, const double*, double*, double*, double*);
    lartg_t dlartg = (lartg_t)dlsym(h, "scipy_dlartg_");
    if (!dlartg) { printf("no dlartg\n"); return 1; }
    std::mt19937 g(5); std::normal_distribution<double> nd;
    double worst = 0;
    std::vector<std::pair<double, double>> cases = {{0, 0}, {1, 0}, {0, 1}, {0, -2}, {-3, 0}, {3, 4}, {-3, 4}, {3, -4}, {-3, -4}, {1e-200, 1e-200}, {1e150, -1e150}};
    for (int i = 0; i < 200; ++i) cases.push_back({nd(g) * pow(10.0, (i % 7) - 3), nd(g)});
    for (auto& fg : cases) {
        double c, s, r, c2, s2, r2;
        lartg(fg.first, fg.second, c, s, r);
        dlartg(&fg.first, &fg.second, &c2, &s2, &r2);
        const double sc = std::max(1.0, fabs(r2));
        worst = std::max(worst, std::max(fabs(c - c2), std::max(fabs(s - s2), fabs(r - r2) / sc)));
    }
    printf("lartg %.3e\n", worst);
    // givens_real: after k steps the rotations triangularise a Hessenberg column sequence: apply to random columns and
    // check (i) the rotation is orthogonal (norm preserved), (ii) h(k+1) is annihilated
    {
        const int kmax = 12; std::vector<cd> c(kmax), s(kmax);
        double wn = 0, wz = 0;
        for (int k = 1; k <= kmax; ++k) {
            std::vector<cd> hcol(k + 1);
            double n0 = 0; for (auto& v : hcol) { v = nd(g); n0 += std::norm(v); }
            givens_real(hcol.data(), c.data(), s.data(), k);
            double n1 = 0; for (auto& v : hcol) n1 += std::norm(v);
            wn = std::max(wn, fabs(sqrt(n1) - sqrt(n0)) / sqrt(n0)); wz = std::max(wz, std::abs(hcol[k]));
            wn = std::max(wn, fabs(std::norm(c[k - 1]) + std::norm(s[k - 1]) - 1.0));
        }
        printf("givens_real norm %.3e zero %.3e\n", wn, wz);
    }
    // givens_cplx on REAL data must agree with the unitary rotation up to the sign convention of r (c, s real: conjugates
    // do not matter); on complex data it is the reference's literal, non-unitary form: |c|^2 + |s|^2 = 1 still holds
    {
        const int kmax = 12; std::vector<cd> c(kmax), s(kmax);
        double wn = 0, wz = 0, wcs = 0;
        for (int k = 1; k <= kmax; ++k) {
            std::vector<cd> hcol(k + 1);
            double n0 = 0; for (auto& v : hcol) { v = nd(g); n0 += std::norm(v); }
            givens_cplx(hcol.data(), c.data(), s.data(), k);
            double n1 = 0; for (auto& v : hcol) n1 += std::norm(v);
            wn = std::max(wn, fabs(sqrt(n1) - sqrt(n0)) / sqrt(n0)); wz = std::max(wz, std::abs(hcol[k]));
        }
        for (int k = 1; k <= kmax; ++k) {
            std::vector<cd> hcol(k + 1);
            for (auto& v : hcol) v = cd(nd(g), nd(g));
            givens_cplx(hcol.data(), c.data(), s.data(), k);
            wcs = std::max(wcs, fabs(std::norm(c[k - 1]) + std::norm(s[k - 1]) - 1.0));
            wz = std::max(wz, std::abs(hcol[k]));
        }
        printf("givens_cplx realdata_norm %.3e zero %.3e cs %.3e\n", wn, wz, wcs);
    }
    return 0;
}
