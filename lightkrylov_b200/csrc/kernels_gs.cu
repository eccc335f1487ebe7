// kernels_gs.cu -- the two HBM-bound kernels of one classical Gram-Schmidt pass.
//
// They replace, on a contiguous column-major device basis,
//   innerprod(X(:j), y)            = j separate dot sweeps   (AbstractVectors.fypp:659-675)
//   y%norm() zero check            = one more sweep           (gram_schmidt.fypp:127)
//   linear_combination + y%sub     = j axpbys on a temporary  (gram_schmidt.fypp:141-146)
// by ONE multi-dot (reads V once, w once) and ONE multi-axpy (reads V once, w once, writes w once).
// Algorithmic bytes per pass: 2*j*n*s (+3*n*s for w).  No tensor cores: AI ~ 0.25 flop/B.
#include "lkb_kernels.h"
#include "lkb_p2p.cuh"

namespace lkb {

// ------------------------------------------------------------------------------------------
// multi-dot.  Persistent grid of 2 CTAs per SM; every CTA owns a contiguous row range and walks
// it in tiles of MD_THREADS*PT packs.  Per tile the w packs are loaded ONCE into registers and
// stay there while the CTA sweeps all j basis columns in chunks of MD_CB: V and w are each read
// exactly once from HBM ((j+1)*n*s bytes per launch).  Every thread keeps MD_CB accumulators
// and issues independent 128-bit loads; at the end of a (tile, chunk) the warp folds its
// MD_CB x 32 partial sums with a transposing butterfly (16 shuffles instead of 80) and adds
// them to its private row of shared-memory accumulators.
// Stage 1: fixed-order sum over the CTA's warps -> one partial row per CTA.  Stage 2: the last
// CTA to retire (atomic ticket) folds the partial rows in fixed order => run-to-run deterministic.
// ------------------------------------------------------------------------------------------
template <typename T> LKB_DI T shfl_xor_t(T v, int m);
template <> LKB_DI float shfl_xor_t<float>(float v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
template <> LKB_DI double shfl_xor_t<double>(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
template <> LKB_DI float2 shfl_xor_t<float2>(float2 v, int m) {
    return make_float2(__shfl_xor_sync(0xffffffffu, v.x, m), __shfl_xor_sync(0xffffffffu, v.y, m));
}
template <> LKB_DI double2 shfl_xor_t<double2>(double2 v, int m) {
    return make_double2(__shfl_xor_sync(0xffffffffu, v.x, m), __shfl_xor_sync(0xffffffffu, v.y, m));
}

// Transposing warp reduction of 16 per-lane values: afterwards acc[0] of lane l holds the warp
// total of value index ((l>>4)&1)*8 + ((l>>3)&1)*4 + ((l>>2)&1)*2 + ((l>>1)&1).  Fixed order.
template <typename E> LKB_DI void warp_fold16(E (&acc)[16], int lane) {
#pragma unroll
    for (int half = 8, bit = 16; half >= 1; half >>= 1, bit >>= 1) {
        const bool up = (lane & bit) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
            const E send = up ? acc[i] : acc[i + half];
            const E keep = up ? acc[i + half] : acc[i];
            acc[i] = add_v(keep, shfl_xor_t<E>(send, bit));
        }
    }
    acc[0] = add_v(acc[0], shfl_xor_t<E>(acc[0], 1));
}

template <int K, int PT>
__global__ void __launch_bounds__(MD_THREADS, 2)
k_multidot(const typename Tr<K>::E* __restrict__ V, int64_t ld, int j,
           const typename Tr<K>::E* __restrict__ w, int64_t n,
           typename Tr<K>::W* __restrict__ partial, typename Tr<K>::W* __restrict__ out,
           unsigned* __restrict__ counter, const int* __restrict__ flags, const P2P p2p)
{
    using E = typename Tr<K>::E;
    using W = typename Tr<K>::W;
    constexpr int EPP = Tr<K>::EPP;
    constexpr int CB = MD_CB;
    // PT = packs of w held in registers per thread (tile = MD_THREADS*PT packs); smaller tiles are
    // chosen for small n so that the tiles still spread evenly over the 2*SM CTAs
    constexpr int NW = MD_THREADS / 32;
    using P = Pack<E, EPP>;
    static_assert(CB == 16, "warp_fold16 assumes 16 columns per chunk");
    if (flags && flags[F_STOP]) return;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    W* sacc = reinterpret_cast<W*>(smem_raw);   // [NW][jp]
    __shared__ bool is_last;
    const int jp = j + 1;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < NW * jp; i += MD_THREADS) sacc[i] = zero_v(W());
    __syncthreads();
    W* myacc = sacc + wid * jp;
    const int fold_idx = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);

    const int64_t npk = n / EPP;
    constexpr int64_t TILE = (int64_t)MD_THREADS * PT;
    const int64_t ntiles = (npk + TILE - 1) / TILE;
    const int64_t tper = (ntiles + gridDim.x - 1) / gridDim.x;
    const int64_t t0 = (int64_t)blockIdx.x * tper;
    const int64_t t1 = min(ntiles, t0 + tper);
    const int nchunk = (j + CB - 1) / CB;
    E accw = zero_v(E());

    for (int64_t t = t0; t < t1; ++t) {
        const int64_t base = t * TILE + threadIdx.x;
        const bool full = (t + 1) * TILE <= npk;
        P wv[PT];
#pragma unroll
        for (int q = 0; q < PT; ++q) {
            const int64_t pk = base + (int64_t)q * MD_THREADS;
            if (full || pk < npk) wv[q] = ld_pack_nc<P>(w + pk * EPP);
            else {
#pragma unroll
                for (int e = 0; e < EPP; ++e) wv[q].v[e] = zero_v(E());
            }
#pragma unroll
            for (int e = 0; e < EPP; ++e) fma_conj(accw, wv[q].v[e], wv[q].v[e]);
        }
        for (int c = 0; c < nchunk; ++c) {
            const int c0 = c * CB;
            const int ncv = min(CB, j - c0);
            const E* vb = V + (int64_t)c0 * ld;
            E acc[CB];
#pragma unroll
            for (int i = 0; i < CB; ++i) acc[i] = zero_v(E());
            if (full && ncv == CB) {
#pragma unroll
                for (int q = 0; q < PT; ++q) {
                    const int64_t off = (base + (int64_t)q * MD_THREADS) * EPP;
                    P v[CB];
#pragma unroll
                    for (int i = 0; i < CB; ++i) v[i] = ld_pack_nc<P>(vb + (int64_t)i * ld + off);
#pragma unroll
                    for (int i = 0; i < CB; ++i)
#pragma unroll
                        for (int e = 0; e < EPP; ++e) fma_conj(acc[i], v[i].v[e], wv[q].v[e]);
                }
            } else {
#pragma unroll
                for (int q = 0; q < PT; ++q) {
                    const int64_t pk = base + (int64_t)q * MD_THREADS;
                    if (pk < npk) {
                        const int64_t off = pk * EPP;
#pragma unroll
                        for (int i = 0; i < CB; ++i) {
                            if (i < ncv) {
                                const P v = ld_pack_nc<P>(vb + (int64_t)i * ld + off);
#pragma unroll
                                for (int e = 0; e < EPP; ++e) fma_conj(acc[i], v.v[e], wv[q].v[e]);
                            }
                        }
                    }
                }
            }
            warp_fold16<E>(acc, lane);
            if ((lane & 1) == 0 && fold_idx < ncv) wadd(myacc[c0 + fold_idx], widen(acc[0]));
        }
    }
    // ragged tail (n not a multiple of the pack width): one thread of CTA 0
    __syncwarp();
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        for (int64_t r = npk * EPP; r < n; ++r) {
            const E wt = w[r];
            for (int i = 0; i < j; ++i) {
                E a = zero_v(E());
                fma_conj(a, V[(int64_t)i * ld + r], wt);
                wadd(myacc[i], widen(a));
            }
            fma_conj(accw, wt, wt);
        }
    }
    {
        W a = warp_sum(widen(accw));
        if (lane == 0) wadd(myacc[j], a);
    }
    // ---- stage 1: fixed-order sum over the warps of this CTA ----
    __syncthreads();
    for (int i = threadIdx.x; i < jp; i += MD_THREADS) {
        W a = sacc[i];
#pragma unroll
        for (int q = 1; q < NW; ++q) wadd(a, sacc[q * jp + i]);
        partial[(int64_t)blockIdx.x * jp + i] = a;
    }
    // ---- stage 2: last CTA folds the partial rows ----
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(counter, 1u) == gridDim.x - 1u);
    __syncthreads();
    if (is_last) {
        __threadfence();
        const int nrb = gridDim.x;
        for (int col = wid; col < jp; col += NW) {
            W a = zero_v(W());
            for (int b = lane; b < nrb; b += 32) wadd(a, __ldcg(&partial[(int64_t)b * jp + col]));
            a = warp_sum(a);
            if (lane == 0) out[col] = a;
        }
        if (threadIdx.x == 0) *counter = 0u;
        if (p2p.world > 1) p2p_allreduce_cta<W>(p2p, out, jp);
    }
}

// ------------------------------------------------------------------------------------------
// Block (blksize > 1) Gram-Schmidt: two right-hand sides per sweep of V.
//   innerprod_matrix / linear_combination_matrix (AbstractVectors.fypp:605-643, 677-695) loop over the p
//   columns of Y, re-reading X for each; here a pair of columns shares one read of every V pack, so a
//   block step with p = 2 moves the same bytes as a single-vector step.  Same tile walk, transposing fold
//   and deterministic two-stage reduction as k_multidot; out / partial are laid out [2][j+1].
template <int K>
__global__ void __launch_bounds__(MD_THREADS, 1)
k_multidot2(const typename Tr<K>::E* __restrict__ V, int64_t ld, int j,
            const typename Tr<K>::E* __restrict__ w0, const typename Tr<K>::E* __restrict__ w1, int64_t n,
            typename Tr<K>::W* __restrict__ partial, typename Tr<K>::W* __restrict__ out,
            unsigned* __restrict__ counter, const int* __restrict__ flags, const P2P p2p)
{
    using E = typename Tr<K>::E;
    using W = typename Tr<K>::W;
    constexpr int EPP = Tr<K>::EPP;
    constexpr int CB = MD_CB, PT = 2, NW = MD_THREADS / 32;
    using P = Pack<E, EPP>;
    if (flags && flags[F_STOP]) return;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    W* sacc = reinterpret_cast<W*>(smem_raw);   // [NW][2][jp]
    __shared__ bool is_last;
    const int jp = j + 1;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < NW * 2 * jp; i += MD_THREADS) sacc[i] = zero_v(W());
    __syncthreads();
    W* myacc = sacc + (size_t)wid * 2 * jp;
    const int fold_idx = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
    const int64_t npk = n / EPP;
    constexpr int64_t TILE = (int64_t)MD_THREADS * PT;
    const int64_t ntiles = (npk + TILE - 1) / TILE;
    const int64_t tper = (ntiles + gridDim.x - 1) / gridDim.x;
    const int64_t t0 = (int64_t)blockIdx.x * tper, t1 = min(ntiles, t0 + tper);
    const int nchunk = (j + CB - 1) / CB;
    E aw0 = zero_v(E()), aw1 = zero_v(E());
    for (int64_t t = t0; t < t1; ++t) {
        const int64_t base = t * TILE + threadIdx.x;
        P u0[PT], u1[PT];
#pragma unroll
        for (int q = 0; q < PT; ++q) {
            const int64_t pk = base + (int64_t)q * MD_THREADS;
            if (pk < npk) { u0[q] = ld_pack_nc<P>(w0 + pk * EPP); u1[q] = ld_pack_nc<P>(w1 + pk * EPP); }
            else {
#pragma unroll
                for (int e = 0; e < EPP; ++e) { u0[q].v[e] = zero_v(E()); u1[q].v[e] = zero_v(E()); }
            }
#pragma unroll
            for (int e = 0; e < EPP; ++e) { fma_conj(aw0, u0[q].v[e], u0[q].v[e]); fma_conj(aw1, u1[q].v[e], u1[q].v[e]); }
        }
        for (int c = 0; c < nchunk; ++c) {
            const int c0 = c * CB;
            const int ncv = min(CB, j - c0);
            const E* vb = V + (int64_t)c0 * ld;
            E a0[CB], a1[CB];
#pragma unroll
            for (int i = 0; i < CB; ++i) { a0[i] = zero_v(E()); a1[i] = zero_v(E()); }
#pragma unroll
            for (int q = 0; q < PT; ++q) {
                const int64_t pk = base + (int64_t)q * MD_THREADS;
                if (pk < npk) {
                    const int64_t off = pk * EPP;
#pragma unroll
                    for (int i = 0; i < CB; ++i) {
                        if (i < ncv) {
                            const P v = ld_pack_nc<P>(vb + (int64_t)i * ld + off);
#pragma unroll
                            for (int e = 0; e < EPP; ++e) { fma_conj(a0[i], v.v[e], u0[q].v[e]); fma_conj(a1[i], v.v[e], u1[q].v[e]); }
                        }
                    }
                }
            }
            warp_fold16<E>(a0, lane);
            warp_fold16<E>(a1, lane);
            if ((lane & 1) == 0 && fold_idx < ncv) {
                wadd(myacc[c0 + fold_idx], widen(a0[0]));
                wadd(myacc[jp + c0 + fold_idx], widen(a1[0]));
            }
        }
    }
    __syncwarp();
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        for (int64_t r = npk * EPP; r < n; ++r) {
            const E x0 = w0[r], x1 = w1[r];
            for (int i = 0; i < j; ++i) {
                E b0 = zero_v(E()), b1 = zero_v(E());
                fma_conj(b0, V[(int64_t)i * ld + r], x0); fma_conj(b1, V[(int64_t)i * ld + r], x1);
                wadd(myacc[i], widen(b0)); wadd(myacc[jp + i], widen(b1));
            }
            fma_conj(aw0, x0, x0); fma_conj(aw1, x1, x1);
        }
    }
    {
        const W b0 = warp_sum(widen(aw0)), b1 = warp_sum(widen(aw1));
        if (lane == 0) { wadd(myacc[j], b0); wadd(myacc[jp + j], b1); }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * jp; i += MD_THREADS) {
        W a = sacc[i];
#pragma unroll
        for (int q = 1; q < NW; ++q) wadd(a, sacc[(size_t)q * 2 * jp + i]);
        partial[(int64_t)blockIdx.x * 2 * jp + i] = a;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(counter, 1u) == gridDim.x - 1u);
    __syncthreads();
    if (is_last) {
        __threadfence();
        const int nrb = gridDim.x;
        for (int col = wid; col < 2 * jp; col += NW) {
            W a = zero_v(W());
            for (int b = lane; b < nrb; b += 32) wadd(a, __ldcg(&partial[(int64_t)b * 2 * jp + col]));
            a = warp_sum(a);
            if (lane == 0) out[col] = a;
        }
        if (threadIdx.x == 0) *counter = 0u;
        if (p2p.world > 1) p2p_allreduce_cta<W>(p2p, out, 2 * jp);
    }
}

// W_q -= V c_q for q = 0, 1 with one read of every V pack; c laid out [2][j+1] (W type).
template <int K>
__global__ void __launch_bounds__(256, 2)
k_multiaxpy2(const typename Tr<K>::E* __restrict__ V, int64_t ld, int j, const typename Tr<K>::W* __restrict__ c,
             typename Tr<K>::E* __restrict__ w0, typename Tr<K>::E* __restrict__ w1, int64_t n, const int* __restrict__ flags)
{
    using E = typename Tr<K>::E;
    constexpr int EPP = Tr<K>::EPP;
    using P = Pack<E, EPP>;
    if (flags && flags[F_STOP]) return;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    E* cs = reinterpret_cast<E*>(smem_raw);      // [2][j]
    for (int i = threadIdx.x; i < j; i += blockDim.x) { narrow(c[i], cs[i]); narrow(c[(j + 1) + i], cs[j + i]); }
    __syncthreads();
    const int64_t npk = n / EPP;
    for (int64_t pk = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; pk < npk; pk += (int64_t)gridDim.x * blockDim.x) {
        const int64_t off = pk * EPP;
        P a0 = ld_pack<P>(w0 + off), a1 = ld_pack<P>(w1 + off);
        const E* vp = V + off;
        int i = 0;
        for (; i + 4 <= j; i += 4) {
            P v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = ld_pack_nc<P>(vp + (int64_t)(i + u) * ld);
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int e = 0; e < EPP; ++e) { fnma(a0.v[e], v[u].v[e], cs[i + u]); fnma(a1.v[e], v[u].v[e], cs[j + i + u]); }
        }
        for (; i < j; ++i) {
            const P v = ld_pack_nc<P>(vp + (int64_t)i * ld);
#pragma unroll
            for (int e = 0; e < EPP; ++e) { fnma(a0.v[e], v.v[e], cs[i]); fnma(a1.v[e], v.v[e], cs[j + i]); }
        }
        st_pack(w0 + off, a0); st_pack(w1 + off, a1);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
        for (int64_t t = npk * EPP; t < n; ++t) {
            E a0 = w0[t], a1 = w1[t];
            for (int i = 0; i < j; ++i) { fnma(a0, V[(int64_t)i * ld + t], cs[i]); fnma(a1, V[(int64_t)i * ld + t], cs[j + i]); }
            w0[t] = a0; w1[t] = a1;
        }
}

// ------------------------------------------------------------------------------------------
// dot of two vectors (the `dot` TBP, CG's p^H Ap, Lanczos' 3-term coefficients): the j = 1 case of the
// multi-dot contract -- out[0] = x^H y, out[1] = y^H y -- as a plain grid-stride kernel with 4 + 4
// independent 128-bit loads in flight per thread (the chunked multi-dot keeps only 1 of its 16 column
// slots busy for j = 1 and ran at 2.4 TB/s).  Same deterministic two-stage reduction + p2p allreduce.
template <int K>
__global__ void __launch_bounds__(256, 4)
k_dot2(const typename Tr<K>::E* __restrict__ x, const typename Tr<K>::E* __restrict__ y, int64_t n,
       typename Tr<K>::W* __restrict__ partial, typename Tr<K>::W* __restrict__ out,
       unsigned* __restrict__ counter, const int* __restrict__ flags, const P2P p2p)
{
    using E = typename Tr<K>::E;
    using W = typename Tr<K>::W;
    constexpr int EPP = Tr<K>::EPP;
    constexpr int U = 4;
    using P = Pack<E, EPP>;
    if (flags && flags[F_STOP]) return;
    const int64_t npk = n / EPP;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    E axy = zero_v(E()), ayy = zero_v(E());
    int64_t pk = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; pk + (U - 1) * stride < npk; pk += U * stride) {
        P xv[U], yv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) { xv[u] = ld_pack_nc<P>(x + (pk + u * stride) * EPP); yv[u] = ld_pack_nc<P>(y + (pk + u * stride) * EPP); }
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int e = 0; e < EPP; ++e) { fma_conj(axy, xv[u].v[e], yv[u].v[e]); fma_conj(ayy, yv[u].v[e], yv[u].v[e]); }
    }
    for (; pk < npk; pk += stride) {
        const P xv = ld_pack_nc<P>(x + pk * EPP), yv = ld_pack_nc<P>(y + pk * EPP);
#pragma unroll
        for (int e = 0; e < EPP; ++e) { fma_conj(axy, xv.v[e], yv.v[e]); fma_conj(ayy, yv.v[e], yv.v[e]); }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
        for (int64_t t = npk * EPP; t < n; ++t) { fma_conj(axy, x[t], y[t]); fma_conj(ayy, y[t], y[t]); }
    __shared__ W sm[8][2];
    __shared__ bool is_last;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const W a = warp_sum(widen(axy)), b = warp_sum(widen(ayy));
    if (lane == 0) { sm[wid][0] = a; sm[wid][1] = b; }
    __syncthreads();
    if (threadIdx.x < 2) {
        W t = sm[0][threadIdx.x];
        for (int q = 1; q < 8; ++q) wadd(t, sm[q][threadIdx.x]);
        partial[(int64_t)blockIdx.x * 2 + threadIdx.x] = t;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(counter, 1u) == gridDim.x - 1u);
    __syncthreads();
    if (is_last) {
        __threadfence();
        if (wid < 2) {
            W t = zero_v(W());
            for (int bq = lane; bq < (int)gridDim.x; bq += 32) wadd(t, __ldcg(&partial[(int64_t)bq * 2 + wid]));
            t = warp_sum(t);
            if (lane == 0) out[wid] = t;
        }
        if (threadIdx.x == 0) *counter = 0u;
        if (p2p.world > 1) p2p_allreduce_cta<W>(p2p, out, 2);
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
            unsigned* __restrict__ counter, const int* __restrict__ flags, const P2P p2p)
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
        if (is_last) {
            if (wid == 0) {
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
            if (p2p.world > 1) p2p_allreduce_cta<W>(p2p, nrm2_out, 1);
        }
    }
}

// ------------------------------------------------------------------------------------------
template <int K>
static void multidot_t(cudaStream_t s, const void* V, int64_t ld, int j, const void* w, int64_t n,
                       void* partial, void* out, unsigned* counter, const int* flags, int sms, const P2P* p2p) {
    using E = typename Tr<K>::E; using W = typename Tr<K>::W;
    if (j == 1) {      // two-vector dot: dedicated streaming kernel
        const int64_t npk1 = n / Tr<K>::EPP;
        int64_t nb1 = (npk1 + 256 * 4 - 1) / (256 * 4);
        if (nb1 < 1) nb1 = 1;
        if (nb1 > 4 * (int64_t)sms) nb1 = 4 * (int64_t)sms;
        if (nb1 > MAX_ROWBLOCKS) nb1 = MAX_ROWBLOCKS;
        k_dot2<K><<<(int)nb1, 256, 0, s>>>((const E*)V, (const E*)w, n, (W*)partial, (W*)out, counter, flags, p2p ? *p2p : P2P());
        return;
    }
    // grid: fixed function of the problem size (2 resident CTAs per SM), never of timing.  Tile size:
    // the largest of 4/2/1 packs per thread that still leaves >= 12 tiles per CTA, so the ceil() in the
    // tile split costs < 8 % (at 1/8 of C2 per GPU, 4-pack tiles gave 3.46 tiles per CTA = 86 % balance).
    const int64_t npk = n / Tr<K>::EPP;
    int64_t nb = 2 * (int64_t)sms;
    int pt = 4;
    while (pt > 1 && (npk + (int64_t)MD_THREADS * pt - 1) / ((int64_t)MD_THREADS * pt) < 12 * nb) pt >>= 1;
    const int64_t ntiles = (npk + (int64_t)MD_THREADS * pt - 1) / ((int64_t)MD_THREADS * pt);
    if (nb > ntiles) nb = ntiles;
    if (nb < 1) nb = 1;
    if (nb > MAX_ROWBLOCKS) nb = MAX_ROWBLOCKS;
    const size_t sh = (size_t)(MD_THREADS / 32) * (size_t)(j + 1) * sizeof(W);
    static const bool attr_once = (cudaFuncSetAttribute(k_multidot<K, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024),
                                   cudaFuncSetAttribute(k_multidot<K, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024),
                                   cudaFuncSetAttribute(k_multidot<K, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024), true);
    (void)attr_once;
    const P2P pp = p2p ? *p2p : P2P();
    if (pt == 4) k_multidot<K, 4><<<(int)nb, MD_THREADS, sh, s>>>((const E*)V, ld, j, (const E*)w, n, (W*)partial, (W*)out, counter, flags, pp);
    else if (pt == 2) k_multidot<K, 2><<<(int)nb, MD_THREADS, sh, s>>>((const E*)V, ld, j, (const E*)w, n, (W*)partial, (W*)out, counter, flags, pp);
    else k_multidot<K, 1><<<(int)nb, MD_THREADS, sh, s>>>((const E*)V, ld, j, (const E*)w, n, (W*)partial, (W*)out, counter, flags, pp);
}
void launch_multidot(int kind, cudaStream_t s, const void* V, int64_t ld, int j, const void* w, int64_t n,
                     void* partial, void* out, unsigned* counter, const int* flags, int sms, const P2P* p2p) {
    switch (kind) {
        case KS: multidot_t<KS>(s, V, ld, j, w, n, partial, out, counter, flags, sms, p2p); break;
        case KD: multidot_t<KD>(s, V, ld, j, w, n, partial, out, counter, flags, sms, p2p); break;
        case KC: multidot_t<KC>(s, V, ld, j, w, n, partial, out, counter, flags, sms, p2p); break;
        default: multidot_t<KZ>(s, V, ld, j, w, n, partial, out, counter, flags, sms, p2p); break;
    }
}

template <int K>
static void multiaxpy_t(cudaStream_t s, const void* V, int64_t ld, int j, const void* c, void* w, int64_t n,
                        bool want_norm, void* partial, void* nrm2_out, unsigned* counter, const int* flags, int sms,
                        const P2P* p2p) {
    using E = typename Tr<K>::E; using W = typename Tr<K>::W;
    const P2P pp = p2p ? *p2p : P2P();
    const int64_t npk = n / Tr<K>::EPP;
    int64_t nb = (npk + 255) / 256;
    if (nb < 1) nb = 1;
    if (nb > 2 * sms * 2) nb = 2 * sms * 2;
    if (nb > MAX_ROWBLOCKS) nb = MAX_ROWBLOCKS;
    const size_t sh = (size_t)(j > 0 ? j : 1) * sizeof(E);
    if (want_norm)
        k_multiaxpy<K, true><<<(int)nb, 256, sh, s>>>((const E*)V, ld, j, (const W*)c, (E*)w, n, (double*)partial, (W*)nrm2_out, counter, flags, pp);
    else
        k_multiaxpy<K, false><<<(int)nb, 256, sh, s>>>((const E*)V, ld, j, (const W*)c, (E*)w, n, (double*)partial, (W*)nrm2_out, counter, flags, pp);
}
void launch_multiaxpy(int kind, cudaStream_t s, const void* V, int64_t ld, int j, const void* c, void* w, int64_t n,
                      bool want_norm, void* partial, void* nrm2_out, unsigned* counter, const int* flags, int sms,
                      const P2P* p2p) {
    switch (kind) {
        case KS: multiaxpy_t<KS>(s, V, ld, j, c, w, n, want_norm, partial, nrm2_out, counter, flags, sms, p2p); break;
        case KD: multiaxpy_t<KD>(s, V, ld, j, c, w, n, want_norm, partial, nrm2_out, counter, flags, sms, p2p); break;
        case KC: multiaxpy_t<KC>(s, V, ld, j, c, w, n, want_norm, partial, nrm2_out, counter, flags, sms, p2p); break;
        default: multiaxpy_t<KZ>(s, V, ld, j, c, w, n, want_norm, partial, nrm2_out, counter, flags, sms, p2p); break;
    }
}

template <int K>
static void multidot2_t(cudaStream_t s, const void* V, int64_t ld, int j, const void* w0, const void* w1, int64_t n,
                        void* partial, void* out, unsigned* counter, const int* flags, int sms, const P2P* p2p) {
    using E = typename Tr<K>::E; using W = typename Tr<K>::W;
    const int64_t npk = n / Tr<K>::EPP;
    const int64_t ntiles = (npk + (int64_t)MD_THREADS * 2 - 1) / ((int64_t)MD_THREADS * 2);
    int64_t nb = sms;
    if (nb > ntiles) nb = ntiles;
    if (nb < 1) nb = 1;
    const size_t sh = (size_t)(MD_THREADS / 32) * 2 * (size_t)(j + 1) * sizeof(W);
    static const bool attr_once = (cudaFuncSetAttribute(k_multidot2<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024), true);
    (void)attr_once;
    k_multidot2<K><<<(int)nb, MD_THREADS, sh, s>>>((const E*)V, ld, j, (const E*)w0, (const E*)w1, n, (W*)partial, (W*)out,
                                                   counter, flags, p2p ? *p2p : P2P());
}
void launch_multidot2(int kind, cudaStream_t s, const void* V, int64_t ld, int j, const void* w0, const void* w1, int64_t n,
                      void* partial, void* out, unsigned* counter, const int* flags, int sms, const P2P* p2p) {
    switch (kind) {
        case KS: multidot2_t<KS>(s, V, ld, j, w0, w1, n, partial, out, counter, flags, sms, p2p); break;
        case KD: multidot2_t<KD>(s, V, ld, j, w0, w1, n, partial, out, counter, flags, sms, p2p); break;
        case KC: multidot2_t<KC>(s, V, ld, j, w0, w1, n, partial, out, counter, flags, sms, p2p); break;
        default: multidot2_t<KZ>(s, V, ld, j, w0, w1, n, partial, out, counter, flags, sms, p2p); break;
    }
}
template <int K>
static void multiaxpy2_t(cudaStream_t s, const void* V, int64_t ld, int j, const void* c, void* w0, void* w1, int64_t n,
                         const int* flags, int sms) {
    using E = typename Tr<K>::E; using W = typename Tr<K>::W;
    int64_t nb = (n / Tr<K>::EPP + 255) / 256;
    if (nb < 1) nb = 1;
    if (nb > 4 * (int64_t)sms) nb = 4 * (int64_t)sms;
    const size_t sh = (size_t)2 * (j > 0 ? j : 1) * sizeof(E);
    static const bool attr_once = (cudaFuncSetAttribute(k_multiaxpy2<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024), true);
    (void)attr_once;
    k_multiaxpy2<K><<<(int)nb, 256, sh, s>>>((const E*)V, ld, j, (const W*)c, (E*)w0, (E*)w1, n, flags);
}
void launch_multiaxpy2(int kind, cudaStream_t s, const void* V, int64_t ld, int j, const void* c, void* w0, void* w1,
                       int64_t n, const int* flags, int sms) {
    switch (kind) {
        case KS: multiaxpy2_t<KS>(s, V, ld, j, c, w0, w1, n, flags, sms); break;
        case KD: multiaxpy2_t<KD>(s, V, ld, j, c, w0, w1, n, flags, sms); break;
        case KC: multiaxpy2_t<KC>(s, V, ld, j, c, w0, w1, n, flags, sms); break;
        default: multiaxpy2_t<KZ>(s, V, ld, j, c, w0, w1, n, flags, sms); break;
    }
}

}  // namespace lkb
