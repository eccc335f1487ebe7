// kernels_gemm.cu -- tall-skinny basis update  Y(:, 0:p) = X(:, 0:k) Z(0:k, 0:p).
//
// This is the only dense contraction adjacent to the hot path: the Krylov-Schur restart
// X(:n) <- X(:kdim) Z(:, :n)  (src/Krylov/BaseKrylov.fypp:816-824) and the Ritz-vector assembly
// X(i) = sum_j y(j,i) Xwrk(j) (IterativeSolvers.fypp:1127-1132, eighs.fypp:116-123,
// svd_solvers.fypp:113-119), which the reference performs as k*p separate axpby sweeps.
// Row-local (no collective).  Each thread owns one 16-byte pack of rows and PB output columns in
// registers; X is streamed once per group of PB outputs with 128-bit loads, Z is broadcast from
// shared memory.  PB = 16 (real) / 8 (complex) keeps the fp64 FMA rate needed at full HBM speed below the
// B200 fp64 pipe, so this stays HBM-bound without tensor cores (a DMMA variant is a later row).
#include "lkb_kernels.h"

namespace lkb {

template <int K, int PB>
__global__ void __launch_bounds__(256)
k_basis_gemm(const typename Tr<K>::E* __restrict__ X, int64_t ldx, int k,
             const typename Tr<K>::E* __restrict__ Z, int ldz, int p,
             typename Tr<K>::E* __restrict__ Y, int64_t ldy, int64_t n)
{
    using E = typename Tr<K>::E;
    constexpr int EPP = Tr<K>::EPP;
    using P = Pack<E, EPP>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    E* zs = reinterpret_cast<E*>(smem_raw);            // [k][PB]
    const int p0 = blockIdx.y * PB;
    const int np = min(PB, p - p0);
    for (int t = threadIdx.x; t < k * PB; t += blockDim.x) {
        const int i = t / PB, q = t % PB;
        zs[t] = q < np ? Z[i + (int64_t)ldz * (p0 + q)] : zero_v(E());
    }
    __syncthreads();
    const int64_t npk = n / EPP;
    for (int64_t pk = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; pk < npk; pk += (int64_t)gridDim.x * blockDim.x) {
        const int64_t off = pk * EPP;
        P acc[PB];
#pragma unroll
        for (int q = 0; q < PB; ++q)
#pragma unroll
            for (int e = 0; e < EPP; ++e) acc[q].v[e] = zero_v(E());
        int i = 0;
        for (; i + 4 <= k; i += 4) {
            P v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = ld_pack_nc<P>(X + (int64_t)(i + u) * ldx + off);
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int q = 0; q < PB; ++q) {
                    const E z = zs[(i + u) * PB + q];
#pragma unroll
                    for (int e = 0; e < EPP; ++e) fmacc(acc[q].v[e], v[u].v[e], z);
                }
        }
        for (; i < k; ++i) {
            const P v = ld_pack_nc<P>(X + (int64_t)i * ldx + off);
#pragma unroll
            for (int q = 0; q < PB; ++q) {
                const E z = zs[i * PB + q];
#pragma unroll
                for (int e = 0; e < EPP; ++e) fmacc(acc[q].v[e], v.v[e], z);
            }
        }
#pragma unroll
        for (int q = 0; q < PB; ++q)
            if (q < np) st_pack(Y + (int64_t)(p0 + q) * ldy + off, acc[q]);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
        for (int64_t r = npk * EPP; r < n; ++r)
            for (int q = 0; q < np; ++q) {
                E a = zero_v(E());
                for (int i = 0; i < k; ++i) fmacc(a, X[(int64_t)i * ldx + r], zs[i * PB + q]);
                Y[(int64_t)(p0 + q) * ldy + r] = a;
            }
}

template <int K>
static void gemm_t(cudaStream_t s, const void* X, int64_t ldx, int k, const void* Z, int ldz, int p, void* Y,
                   int64_t ldy, int64_t n, int sms) {
    using E = typename Tr<K>::E;
    // output columns per sweep of X: 16 for the real kinds (2 * 16 FMA per 16-byte pack = 26 TFLOP/s fp64 at
    // full HBM speed, ~70 % of the B200 fp64 pipe), 8 for the complex kinds (4 FMA per complex multiply-add)
    constexpr int PB = Tr<K>::cplx ? 8 : 16;
    const int64_t npk = n / Tr<K>::EPP;
    int64_t nb = (npk + 255) / 256;
    if (nb < 1) nb = 1;
    if (nb > 4 * (int64_t)sms) nb = 4 * (int64_t)sms;
    dim3 grid((unsigned)nb, (unsigned)((p + PB - 1) / PB));
    const size_t sh = (size_t)k * PB * sizeof(E);
    static const bool attr_once = (cudaFuncSetAttribute(k_basis_gemm<K, PB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024), true);
    (void)attr_once;
    k_basis_gemm<K, PB><<<grid, 256, sh, s>>>((const E*)X, ldx, k, (const E*)Z, ldz, p, (E*)Y, ldy, n);
}
void launch_basis_gemm(int kind, cudaStream_t s, const void* X, int64_t ldx, int k, const void* Z, int ldz, int p,
                       void* Y, int64_t ldy, int64_t n, int sms) {
    switch (kind) {
        case KS: gemm_t<KS>(s, X, ldx, k, Z, ldz, p, Y, ldy, n, sms); break;
        case KD: gemm_t<KD>(s, X, ldx, k, Z, ldz, p, Y, ldy, n, sms); break;
        case KC: gemm_t<KC>(s, X, ldx, k, Z, ldz, p, Y, ldy, n, sms); break;
        default: gemm_t<KZ>(s, X, ldx, k, Z, ldz, p, Y, ldy, n, sms); break;
    }
}

}  // namespace lkb
