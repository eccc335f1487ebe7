// kernels_gs.cu -- the two HBM-bound kernels of one classical Gram-Schmidt pass.
//
// They replace, on a contiguous column-major device basis,
//   innerprod(X(:j), y)            = j separate dot sweeps   (AbstractVectors.fypp:659-675)
//   y%norm() zero check            = one more sweep           (gram_schmidt.fypp:127)
//   linear_combination + y%sub     = j axpbys on a temporary  (gram_schmidt.fypp:141-146)
// by ONE multi-dot (reads V once, w once) and ONE multi-axpy (reads V once, w once, writes w once).
// Algorithmic bytes per pass: 2*j*n*s (+3*n*s for w).  No tensor cores: AI ~ 0.25 flop/B.
#include "lkb_kernels.h"

namespace lkb {

// ------------------------------------------------------------------------------------------
// multi-dot:  grid = (nchunks, nrb).  blockIdx.x (fastest-scheduled) selects a chunk of MD_CB
// basis columns, blockIdx.y a contiguous row range, so the CTAs that share a row range of w are
// co-resident and w is served from L2 after its first read.  Every thread keeps MD_CB
// accumulators and issues MD_CB independent 128-bit loads per iteration.
// Stage 1: warp shuffle -> smem -> one partial row per (row block).  Stage 2: the last CTA to
// retire (atomic ticket) folds the nrb partial rows in a fixed order => run-to-run deterministic.
// ------------------------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(MD_THREADS, 2)
k_multidot(const typename Tr<K>::E* __restrict__ V, int64_t ld, int j,
           const typename Tr<K>::E* __restrict__ w, int64_t n,
           typename Tr<K>::W* __restrict__ partial, typename Tr<K>::W* __restrict__ out,
           unsigned* __restrict__ counter, const int* __restrict__ flags)
{
    using E = typename Tr<K>::E;
    using W = typename Tr<K>::W;
    constexpr int EPP = Tr<K>::EPP;
    constexpr int CB = MD_CB;
    using P = Pack<E, EPP>;
    if (flags && flags[F_STOP]) return;

    const int jp = j + 1;
    const int c0 = blockIdx.x * CB;
    const int ncv = min(CB, j - c0) > 0 ? min(CB, j - c0) : 0;   // basis columns in this chunk
    const bool do_ww = (blockIdx.x == 0);
    const int64_t npk = n / EPP;
    const int64_t per = (npk + gridDim.y - 1) / gridDim.y;
    const int64_t p0 = (int64_t)blockIdx.y * per;
    const int64_t p1 = min(npk, p0 + per);
    const E* vb = V + (int64_t)c0 * ld;

    E acc[CB];
#pragma unroll
    for (int i = 0; i < CB; ++i) acc[i] = zero_v(E());
    E accw = zero_v(E());

    if (ncv == CB) {
        for (int64_t pk = p0 + threadIdx.x; pk < p1; pk += MD_THREADS) {
            const int64_t off = pk * EPP;
            const P wv = ld_pack_nc<P>(w + off);
            P v[CB];
#pragma unroll
            for (int i = 0; i < CB; ++i) v[i] = ld_pack_nc<P>(vb + (int64_t)i * ld + off);
#pragma unroll
            for (int i = 0; i < CB; ++i)
#pragma unroll
                for (int e = 0; e < EPP; ++e) fma_conj(acc[i], v[i].v[e], wv.v[e]);
            if (do_ww)
#pragma unroll
                for (int e = 0; e < EPP; ++e) fma_conj(accw, wv.v[e], wv.v[e]);
        }
    } else {
        for (int64_t pk = p0 + threadIdx.x; pk < p1; pk += MD_THREADS) {
            const int64_t off = pk * EPP;
            const P wv = ld_pack_nc<P>(w + off);
#pragma unroll
            for (int i = 0; i < CB; ++i) {
                if (i < ncv) {
                    const P v = ld_pack_nc<P>(vb + (int64_t)i * ld + off);
#pragma unroll
                    for (int e = 0; e < EPP; ++e) fma_conj(acc[i], v.v[e], wv.v[e]);
                }
            }
            if (do_ww)
#pragma unroll
                for (int e = 0; e < EPP; ++e) fma_conj(accw, wv.v[e], wv.v[e]);
        }
    }
    // ragged tail (n not a multiple of the pack width): one thread of row block 0
    if (blockIdx.y == 0 && threadIdx.x == 0) {
        for (int64_t t = npk * EPP; t < n; ++t) {
            const E wt = w[t];
#pragma unroll
            for (int i = 0; i < CB; ++i)
                if (i < ncv) fma_conj(acc[i], vb[(int64_t)i * ld + t], wt);
            if (do_ww) fma_conj(accw, wt, wt);
        }
    }

    // ---- stage 1: CTA reduction in fixed order ----
    __shared__ W sm[MD_THREADS / 32][CB + 1];
    __shared__ bool is_last;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < CB; ++i) {
        W a = warp_sum(widen(acc[i]));
        if (lane == 0) sm[wid][i] = a;
    }
    {
        W a = warp_sum(widen(accw));
        if (lane == 0) sm[wid][CB] = a;
    }
    __syncthreads();
    if (threadIdx.x <= CB) {
        const int i = threadIdx.x;
        W a = sm[0][i];
#pragma unroll
        for (int q = 1; q < MD_THREADS / 32; ++q) wadd(a, sm[q][i]);
        if (i < ncv) partial[(int64_t)blockIdx.y * jp + c0 + i] = a;
        else if (i == CB && do_ww) partial[(int64_t)blockIdx.y * jp + j] = a;
    }
    // ---- stage 2: last CTA folds the partial rows ----
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned total = gridDim.x * gridDim.y;
        is_last = (atomicAdd(counter, 1u) == total - 1u);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        const int nrb = gridDim.y;
        for (int col = wid; col < jp; col += MD_THREADS / 32) {
            W a = zero_v(W());
            for (int b = lane; b < nrb; b += 32) wadd(a, __ldcg(&partial[(int64_t)b * jp + col]));
            a = warp_sum(a);
            if (lane == 0) out[col] = a;
        }
        if (threadIdx.x == 0) *counter = 0u;
    }
}

// ------------------------------------------------------------------------------------------
// multi-axpy:  w -= V(:, 0:j) c.  One 16-byte pack of w per thread per iteration, the j basis
// packs streamed with UA independent 128-bit loads in flight; c broadcast from shared memory.
// Optional epilogue: ||w_new||^2 with the same two-stage deterministic reduction (this is the
// norm the reference recomputes in the next pass' zero check / in qr_no_pivoting).
// ------------------------------------------------------------------------------------------
template <int K, bool NORM>
__global__ void __launch_bounds__(256, 2)
k_multiaxpy(const typename Tr<K>::E* __restrict__ V, int64_t ld, int j,
            const typename Tr<K>::W* __restrict__ c, typename Tr<K>::E* __restrict__ w, int64_t n,
            double* __restrict__ partial, typename Tr<K>::W* __restrict__ nrm2_out,
            unsigned* __restrict__ counter, const int* __restrict__ flags)
{
    using E = typename Tr<K>::E;
    using W = typename Tr<K>::W;
    constexpr int EPP = Tr<K>::EPP;
    constexpr int UA = 8;
    using P = Pack<E, EPP>;
    if (flags && flags[F_STOP]) return;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    E* cs = reinterpret_cast<E*>(smem_raw);
    for (int i = threadIdx.x; i < j; i += blockDim.x) narrow(c[i], cs[i]);
    __syncthreads();

    const int64_t npk = n / EPP;
    double nrm = 0.0;
    for (int64_t pk = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; pk < npk; pk += (int64_t)gridDim.x * blockDim.x) {
        const int64_t off = pk * EPP;
        P a = ld_pack<P>(w + off);
        const E* vp = V + off;
        int i = 0;
        for (; i + UA <= j; i += UA) {
            P v[UA];
#pragma unroll
            for (int u = 0; u < UA; ++u) v[u] = ld_pack_nc<P>(vp + (int64_t)(i + u) * ld);
#pragma unroll
            for (int u = 0; u < UA; ++u) {
                const E ci = cs[i + u];
#pragma unroll
                for (int e = 0; e < EPP; ++e) fnma(a.v[e], v[u].v[e], ci);
            }
        }
        for (; i < j; ++i) {
            const P v = ld_pack_nc<P>(vp + (int64_t)i * ld);
            const E ci = cs[i];
#pragma unroll
            for (int e = 0; e < EPP; ++e) fnma(a.v[e], v.v[e], ci);
        }
        st_pack(w + off, a);
        if (NORM)
#pragma unroll
            for (int e = 0; e < EPP; ++e) nrm += abs2_w(a.v[e]);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        for (int64_t t = npk * EPP; t < n; ++t) {
            E a = w[t];
            for (int i = 0; i < j; ++i) fnma(a, V[(int64_t)i * ld + t], cs[i]);
            w[t] = a;
            if (NORM) nrm += abs2_w(a);
        }
    }
    if (NORM) {
        __shared__ double sm[8];
        __shared__ bool is_last;
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        double a = warp_sum(nrm);
        if (lane == 0) sm[wid] = a;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = sm[0];
            for (int q = 1; q < (int)(blockDim.x >> 5); ++q) t += sm[q];
            partial[blockIdx.x] = t;
            __threadfence();
            is_last = (atomicAdd(counter, 1u) == gridDim.x - 1u);
        }
        __syncthreads();
        if (is_last && wid == 0) {
            __threadfence();
            double t = 0.0;
            for (int b = lane; b < (int)gridDim.x; b += 32) t += __ldcg(&partial[b]);
            t = warp_sum(t);
            if (lane == 0) {
                W o = zero_v(W());
                *reinterpret_cast<double*>(&o) = t;   // real part
                nrm2_out[0] = o;
                *counter = 0u;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
static inline int grid_rowblocks(int64_t npk, int nchunks, int sms) {
    // Fixed function of the problem (not of timing): 2 resident CTAs per SM, ~2 waves.
    int target = (2 * sms * 2 + nchunks - 1) / nchunks;
    if (target < 1) target = 1;
    int64_t maxrb = (npk + MD_THREADS - 1) / MD_THREADS;   // at least one pack per thread
    if (maxrb < 1) maxrb = 1;
    if (target > maxrb) target = (int)maxrb;
    if (target > MAX_ROWBLOCKS) target = MAX_ROWBLOCKS;
    return target;
}

template <int K>
static void multidot_t(cudaStream_t s, const void* V, int64_t ld, int j, const void* w, int64_t n,
                       void* partial, void* out, unsigned* counter, const int* flags, int sms) {
    using E = typename Tr<K>::E; using W = typename Tr<K>::W;
    const int nchunks = j > 0 ? (j + MD_CB - 1) / MD_CB : 1;
    const int nrb = grid_rowblocks(n / Tr<K>::EPP, nchunks, sms);
    dim3 grid(nchunks, nrb);
    k_multidot<K><<<grid, MD_THREADS, 0, s>>>((const E*)V, ld, j, (const E*)w, n, (W*)partial, (W*)out, counter, flags);
}
void launch_multidot(int kind, cudaStream_t s, const void* V, int64_t ld, int j, const void* w, int64_t n,
                     void* partial, void* out, unsigned* counter, const int* flags, int sms) {
    switch (kind) {
        case KS: multidot_t<KS>(s, V, ld, j, w, n, partial, out, counter, flags, sms); break;
        case KD: multidot_t<KD>(s, V, ld, j, w, n, partial, out, counter, flags, sms); break;
        case KC: multidot_t<KC>(s, V, ld, j, w, n, partial, out, counter, flags, sms); break;
        default: multidot_t<KZ>(s, V, ld, j, w, n, partial, out, counter, flags, sms); break;
    }
}

template <int K>
static void multiaxpy_t(cudaStream_t s, const void* V, int64_t ld, int j, const void* c, void* w, int64_t n,
                        bool want_norm, void* partial, void* nrm2_out, unsigned* counter, const int* flags, int sms) {
    using E = typename Tr<K>::E; using W = typename Tr<K>::W;
    const int64_t npk = n / Tr<K>::EPP;
    int64_t nb = (npk + 255) / 256;
    if (nb < 1) nb = 1;
    if (nb > 2 * sms * 2) nb = 2 * sms * 2;
    if (nb > MAX_ROWBLOCKS) nb = MAX_ROWBLOCKS;
    const size_t sh = (size_t)(j > 0 ? j : 1) * sizeof(E);
    if (want_norm)
        k_multiaxpy<K, true><<<(int)nb, 256, sh, s>>>((const E*)V, ld, j, (const W*)c, (E*)w, n, (double*)partial, (W*)nrm2_out, counter, flags);
    else
        k_multiaxpy<K, false><<<(int)nb, 256, sh, s>>>((const E*)V, ld, j, (const W*)c, (E*)w, n, (double*)partial, (W*)nrm2_out, counter, flags);
}
void launch_multiaxpy(int kind, cudaStream_t s, const void* V, int64_t ld, int j, const void* c, void* w, int64_t n,
                      bool want_norm, void* partial, void* nrm2_out, unsigned* counter, const int* flags, int sms) {
    switch (kind) {
        case KS: multiaxpy_t<KS>(s, V, ld, j, c, w, n, want_norm, partial, nrm2_out, counter, flags, sms); break;
        case KD: multiaxpy_t<KD>(s, V, ld, j, c, w, n, want_norm, partial, nrm2_out, counter, flags, sms); break;
        case KC: multiaxpy_t<KC>(s, V, ld, j, c, w, n, want_norm, partial, nrm2_out, counter, flags, sms); break;
        default: multiaxpy_t<KZ>(s, V, ld, j, c, w, n, want_norm, partial, nrm2_out, counter, flags, sms); break;
    }
}

}  // namespace lkb
