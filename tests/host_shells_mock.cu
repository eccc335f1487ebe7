// host_shells_mock.cu -- TEST INFRASTRUCTURE: runs the product's host shells (lkb_eig.cu: eigs, eighs, svds, krylov_schur --
// their control flow, LAPACK calls, sorting, restarts, post-processing, write_intermediate side effects) WITHOUT a GPU.
//
// The translation unit lkb_eig.cu is included verbatim.  Everything it reaches on the device side is replaced here:
//   * the CUDA runtime calls it makes directly are redirected to host equivalents by macros ("device" memory = host memory);
//   * the internal step API (arnoldi / lanczos / bidiag enqueue-fetch-collect, basis GEMM, vector helpers, allocation) is
//     implemented on the CPU -- the Krylov steps by the C oracle (oracle/liblk_oracle.so), so that the Hessenberg / tridiagonal /
//     bidiagonal columns the shells see are exactly those the Python oracle shells see.
// tests/test_host_shells_mock.py then requires the C++ shells and the Python oracle shells to agree.  Nothing here is part of
// the product: the library never contains this file, and the GPU parity tests remain the check of the real device path.
#include <complex>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <vector>
#include "../include/lkb.h"
#include "lkb_internal.h"

// ---- CUDA runtime calls made directly by lkb_eig.cu -> host ----------------------------------------------------------------
static cudaError_t mock_malloc_host(void** p, size_t n) { *p = calloc(1, n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
static cudaError_t mock_memcpy2d(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height) {
    for (size_t r = 0; r < height; ++r) memcpy((char*)dst + r * dpitch, (const char*)src + r * spitch, width);
    return cudaSuccess;
}
#define cudaMallocHost(p, n) mock_malloc_host((void**)(p), (n))
#define cudaFreeHost(p) (free(p), cudaSuccess)
#define cudaEventCreateWithFlags(e, f) (*(e) = (cudaEvent_t)(uintptr_t)1, cudaSuccess)
#define cudaEventDestroy(e) ((void)(e), cudaSuccess)
#define cudaEventRecord(e, s) ((void)(e), (void)(s), cudaSuccess)
#define cudaEventSynchronize(e) ((void)(e), cudaSuccess)
#define cudaStreamSynchronize(s) ((void)(s), cudaSuccess)
#define cudaMemcpyAsync(d, s, n, k, st) (memmove((d), (s), (n)), (void)(st), cudaSuccess)
#define cudaMemcpy2DAsync(d, dp, s, sp, w, h, k, st) mock_memcpy2d((d), (dp), (s), (sp), (w), (h))

#include "lkb_eig.cu"        // the product's host shells, verbatim

// ---- the C oracle (oracle/lk_oracle.c): one Krylov step at a time -----------------------------------------------------------
extern "C" {
#define ORACLE_DECL(sfx, R)                                                                                                   \
    int lko_arnoldi_##sfx(void* A, int64_t n, void* X, int64_t ldx, void* H, int ldh, int kdim, int kstart, int kend, R tol,   \
                          int trans, int p, uint64_t* seed);                                                                  \
    int lko_lanczos_##sfx(void* A, int64_t n, void* X, int64_t ldx, void* T, int ldt, int kdim, int kstart, int kend, R tol);   \
    int lko_bidiag_##sfx(void* A, int64_t m, int64_t n, void* U, int64_t ldu, void* V, int64_t ldv, void* B, int ldb, int kdim, \
                         int kstart, int kend, R tol);                                                                        \
    void lko_fill_##sfx(int64_t n, void* x, int dist, uint64_t seed, int64_t row0);
ORACLE_DECL(s, float) ORACLE_DECL(d, double) ORACLE_DECL(c, float) ORACLE_DECL(z, double)
#undef ORACLE_DECL
}

namespace {
char g_err[512] = "";
std::vector<char> g_col;          // the H / T / B column of the step enqueued last (one step is in flight at a time)
int g_info = 0;
uint64_t g_oracle_seed = 1000;
uint64_t g_uid = 0;

template <typename F> void dispatch(int kind, F f) {
    switch (kind) {
        case KS: f((float*)nullptr); break;
        case KD: f((double*)nullptr); break;
        case KC: f((std::complex<float>*)nullptr); break;
        default: f((std::complex<double>*)nullptr); break;
    }
}
template <typename E> E mk_elem(Scalar s);
template <> float mk_elem<float>(Scalar s) { return (float)s.re; }
template <> double mk_elem<double>(Scalar s) { return s.re; }
template <> std::complex<float> mk_elem<std::complex<float>>(Scalar s) { return {(float)s.re, (float)s.im}; }
template <> std::complex<double> mk_elem<std::complex<double>>(Scalar s) { return {s.re, s.im}; }

void stash_column(const void* H, int ldh, int col, size_t es) {
    g_col.assign((size_t)ldh * es, 0);
    memcpy(g_col.data(), (const char*)H + (size_t)ldh * col * es, (size_t)ldh * es);
}
}  // namespace

namespace lkb {

void set_error(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap); }
int bcast_host(lkb_ctx_s*, void*, size_t) { return 0; }
void prof_begin(lkb_ctx_s*, int) {}
void prof_end(lkb_ctx_s*, int, int) {}
int check_launch(lkb_ctx_s*, const char*) { return 0; }
uint64_t next_seed(lkb_ctx_s* c) { return c->seed + 0x9E3779B97F4A7C15ULL * (++c->seed_calls); }
int dev_alloc(lkb_ctx_s*, void** p, size_t bytes) { *p = calloc(1, bytes ? bytes : 16); return *p ? 0 : LKB_ERR_ALLOC; }
void dev_free(lkb_ctx_s*, void* p) { free(p); }
int ensure_hstage(lkb_ctx_s* c, size_t bytes) {
    if (c->hstage_bytes >= bytes) return 0;
    free(c->hstage); c->hstage = calloc(1, bytes); c->hstage_bytes = bytes;
    return 0;
}
int ensure_coefd(lkb_ctx_s* c, size_t bytes) {
    if (c->coefd_bytes >= bytes) return 0;
    free(c->coefd); c->coefd = calloc(1, bytes); c->coefd_bytes = bytes;
    return 0;
}

void launch_fill(int kind, cudaStream_t, void* x, int64_t n, int64_t row0, int dist, uint64_t seed, int) {
    switch (kind) {
        case KS: lko_fill_s(n, x, dist, seed, row0); break;
        case KD: lko_fill_d(n, x, dist, seed, row0); break;
        case KC: lko_fill_c(n, x, dist, seed, row0); break;
        default: lko_fill_z(n, x, dist, seed, row0); break;
    }
}
void launch_scal(int kind, cudaStream_t, Scalar alpha, void* x, int64_t n, int) {
    dispatch(kind, [&](auto* tag) {
        typedef typename std::remove_pointer<decltype(tag)>::type E;
        const E a = mk_elem<E>(alpha); E* v = (E*)x;
        for (int64_t i = 0; i < n; ++i) v[i] = a * v[i];
    });
}
void launch_axpby(int kind, cudaStream_t, Scalar alpha, const void* x, Scalar beta, void* y, int64_t n, int) {
    dispatch(kind, [&](auto* tag) {
        typedef typename std::remove_pointer<decltype(tag)>::type E;
        const E a = mk_elem<E>(alpha), b = mk_elem<E>(beta); const E* u = (const E*)x; E* v = (E*)y;
        const bool overwrite = (beta.re == 0.0 && beta.im == 0.0);                 // beta == 0: y is not read (copy)
        for (int64_t i = 0; i < n; ++i) v[i] = overwrite ? a * u[i] : a * u[i] + b * v[i];
    });
}
int vec_norm_sync(lkb_ctx_s*, int kind, const void* w, int64_t n, double* out) {
    double s = 0.0;
    dispatch(kind, [&](auto* tag) {
        typedef typename std::remove_pointer<decltype(tag)>::type E;
        const E* v = (const E*)w;
        for (int64_t i = 0; i < n; ++i) s += std::norm(std::complex<double>(v[i]));
    });
    *out = sqrt(s);
    return 0;
}
// Y(:, q) = sum_i X(:, i) Z(i, q)
void launch_basis_gemm(int kind, cudaStream_t, const void* X, int64_t ldx, int k, const void* Z, int ldz, int p, void* Y, int64_t ldy,
                       int64_t n, int) {
    dispatch(kind, [&](auto* tag) {
        typedef typename std::remove_pointer<decltype(tag)>::type E;
        const E* x = (const E*)X; const E* z = (const E*)Z; E* y = (E*)Y;
        for (int q = 0; q < p; ++q)
            for (int64_t r = 0; r < n; ++r) {
                std::complex<double> acc = 0;
                for (int i = 0; i < k; ++i) acc += std::complex<double>(x[r + ldx * i]) * std::complex<double>(z[i + (int64_t)ldz * q]);
                if constexpr (std::is_floating_point<E>::value) y[r + ldy * q] = (E)acc.real();
                else y[r + ldy * q] = E((typename E::value_type)acc.real(), (typename E::value_type)acc.imag());
            }
    });
}

// ---- Krylov steps: one step of the C oracle on the host copy of the basis; the new column travels through g_col -----------------
int arnoldi_enqueue(lkb_op_s* A, lkb_basis_s* X, int kstart, int kend, double tol, bool tr) {
    if (kstart != kend) { set_error("mock: one step at a time"); return LKB_ERR_ARG; }
    const int kdim = X->ncols - 1, ldh = kdim + 1; const size_t es = kind_size(X->kind);
    std::vector<char> H((size_t)ldh * kdim * es, 0);
    switch (X->kind) {
        case KS: g_info = lko_arnoldi_s(A->user, X->n, X->d, X->ld, H.data(), ldh, kdim, kstart, kend, (float)tol, tr, 1, &g_oracle_seed); break;
        case KD: g_info = lko_arnoldi_d(A->user, X->n, X->d, X->ld, H.data(), ldh, kdim, kstart, kend, tol, tr, 1, &g_oracle_seed); break;
        case KC: g_info = lko_arnoldi_c(A->user, X->n, X->d, X->ld, H.data(), ldh, kdim, kstart, kend, (float)tol, tr, 1, &g_oracle_seed); break;
        default: g_info = lko_arnoldi_z(A->user, X->n, X->d, X->ld, H.data(), ldh, kdim, kstart, kend, tol, tr, 1, &g_oracle_seed); break;
    }
    stash_column(H.data(), ldh, kstart - 1, es);
    return 0;
}
static void fetch_into(void* slot, int ld, size_t es) {
    memcpy(slot, g_col.data(), (size_t)ld * es);
    memcpy((char*)slot + (size_t)ld * es, &g_info, sizeof(int));
}
static void collect_from(const void* slot, void* M, int ld, int k, size_t es, int32_t* info) {
    memcpy((char*)M + (size_t)ld * (k - 1) * es, slot, (size_t)std::min(ld, k + 1) * es);       // rows 0..k of column k
    int v; memcpy(&v, (const char*)slot + (size_t)ld * es, sizeof(int)); *info = v;
}
int arnoldi_fetch_async(lkb_basis_s* X, int, int, void* slot) { fetch_into(slot, X->ncols, kind_size(X->kind)); return 0; }
int arnoldi_collect(lkb_op_s* A, lkb_basis_s* X, void* H, int ldh, int32_t* info, int kstart, int, bool tr, const void* slot) {
    if (ldh != X->ncols) { set_error("mock: ldh"); return LKB_ERR_ARG; }
    collect_from(slot, H, ldh, kstart, kind_size(X->kind), info);
    if (tr) A->n_rmatvec++; else A->n_matvec++;
    return 0;
}
int lanczos_enqueue(lkb_op_s* A, lkb_basis_s* X, int kstart, int kend, double tol) {
    if (kstart != kend) { set_error("mock: one step at a time"); return LKB_ERR_ARG; }
    const int kdim = X->ncols - 1, ldt = kdim + 1; const size_t es = kind_size(X->kind);
    std::vector<char> T((size_t)ldt * kdim * es, 0);
    switch (X->kind) {
        case KS: g_info = lko_lanczos_s(A->user, X->n, X->d, X->ld, T.data(), ldt, kdim, kstart, kend, (float)tol); break;
        case KD: g_info = lko_lanczos_d(A->user, X->n, X->d, X->ld, T.data(), ldt, kdim, kstart, kend, tol); break;
        case KC: g_info = lko_lanczos_c(A->user, X->n, X->d, X->ld, T.data(), ldt, kdim, kstart, kend, (float)tol); break;
        default: g_info = lko_lanczos_z(A->user, X->n, X->d, X->ld, T.data(), ldt, kdim, kstart, kend, tol); break;
    }
    stash_column(T.data(), ldt, kstart - 1, es);
    return 0;
}
int krylov_fetch_async(lkb_ctx_s*, int kind, int ld, int, int, void* slot) { fetch_into(slot, ld, kind_size(kind)); return 0; }
int lanczos_collect(lkb_op_s* A, lkb_basis_s* X, void* T, int ldt, int32_t* info, int kstart, int, const void* slot) {
    collect_from(slot, T, ldt, kstart, kind_size(X->kind), info);
    A->n_matvec++;
    return 0;
}
static lkb_basis_s* g_V = nullptr;     // bidiag_collect does not receive V: remembered from bidiag_enqueue
int bidiag_enqueue(lkb_op_s* A, lkb_basis_s* U, lkb_basis_s* V, int kstart, int kend, double tol) {
    if (kstart != kend) { set_error("mock: one step at a time"); return LKB_ERR_ARG; }
    const int kdim = U->ncols - 1, ldb = kdim + 1; const size_t es = kind_size(U->kind);
    std::vector<char> B((size_t)ldb * kdim * es, 0);
    g_V = V;
    switch (U->kind) {
        case KS: g_info = lko_bidiag_s(A->user, U->n, V->n, U->d, U->ld, V->d, V->ld, B.data(), ldb, kdim, kstart, kend, (float)tol); break;
        case KD: g_info = lko_bidiag_d(A->user, U->n, V->n, U->d, U->ld, V->d, V->ld, B.data(), ldb, kdim, kstart, kend, tol); break;
        case KC: g_info = lko_bidiag_c(A->user, U->n, V->n, U->d, U->ld, V->d, V->ld, B.data(), ldb, kdim, kstart, kend, (float)tol); break;
        default: g_info = lko_bidiag_z(A->user, U->n, V->n, U->d, U->ld, V->d, V->ld, B.data(), ldb, kdim, kstart, kend, tol); break;
    }
    stash_column(B.data(), ldb, kstart - 1, es);
    return 0;
}
int bidiag_collect(lkb_op_s* A, lkb_basis_s* U, void* B, int ldb, int32_t* info, int kstart, int, const void* slot) {
    collect_from(slot, B, ldb, kstart, kind_size(U->kind), info);
    A->n_matvec++; A->n_rmatvec++;
    return 0;
}

}  // namespace lkb

// ---- the few C-ABI entry points lkb_eig.cu calls, on host memory ------------------------------------------------------------------
extern "C" {

int lkb_basis_create(lkb_ctx_t c, int kind, int64_t n_local, int64_t n_global, int64_t row0, int ncols, lkb_basis_t* b) {
    const int64_t ld = std::max<int64_t>(n_local, 1);
    lkb_basis_s* h = new lkb_basis_s{c, kind, n_local, n_global, row0, ld, ncols, nullptr, ++g_uid};
    h->d = calloc((size_t)ld * ncols, kind_size(kind));
    *b = h;
    return 0;
}
int lkb_basis_destroy(lkb_basis_t b) { if (!b) return LKB_ERR_ARG; if (b->owns) free(b->d); delete b; return 0; }
int lkb_basis_zero(lkb_basis_t b, int col0, int ncols) {
    if (col0 < 0 || col0 + ncols > b->ncols) return LKB_ERR_ARG;
    memset(col_ptr(b, col0), 0, (size_t)b->ld * ncols * kind_size(b->kind));
    return 0;
}
int lkb_sync(lkb_ctx_t) { return 0; }

// ---- harness API for tests/test_host_shells_mock.py ---------------------------------------------------------------------------------
const char* mock_last_error(void) { return g_err; }
void* mock_ctx_new(int write_intermediate) { lkb_ctx_s* c = new lkb_ctx_s(); c->write_intermediate = write_intermediate != 0; return c; }
void mock_ctx_free(void* c) { lkb_ctx_s* x = (lkb_ctx_s*)c; free(x->hstage); free(x->coefd); delete x; }
// operator = a pointer to the oracle's lko_op struct of the same kind (kept alive by the caller)
void* mock_op_new(void* ctx, int kind, int64_t m, int64_t n, void* oracle_op) {
    lkb_op_s* A = new lkb_op_s();
    A->ctx = (lkb_ctx_s*)ctx; A->type = 9; A->kind = kind; A->m = m; A->n = n; A->user = oracle_op;
    return A;
}
void mock_op_free(void* A) { delete (lkb_op_s*)A; }
void* mock_vec_new(void* ctx, int kind, int64_t n, const void* host) {
    lkb_vec_s* v = new lkb_vec_s{(lkb_ctx_s*)ctx, kind, n, n, 0, nullptr, true};
    v->d = malloc((size_t)std::max<int64_t>(n, 1) * kind_size(kind));
    memcpy(v->d, host, (size_t)n * kind_size(kind));
    return v;
}
void mock_vec_free(void* v) { free(((lkb_vec_s*)v)->d); delete (lkb_vec_s*)v; }
void mock_basis_get(void* b, void* host) { lkb_basis_s* B = (lkb_basis_s*)b; memcpy(host, B->d, (size_t)B->ld * B->ncols * kind_size(B->kind)); }
void mock_basis_put(void* b, const void* host) { lkb_basis_s* B = (lkb_basis_s*)b; memcpy(B->d, host, (size_t)B->ld * B->ncols * kind_size(B->kind)); }
void mock_set_oracle_seed(uint64_t s) { g_oracle_seed = s; }

}  // extern "C"
