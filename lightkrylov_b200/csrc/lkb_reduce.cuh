// lkb_reduce.cuh -- second stage of the deterministic two-stage reductions.
//
// Stage 1 leaves one row of jp partial sums per CTA in `partial` ([CTA][jp], W type).  Round 1 let the LAST CTA to
// retire fold all rows alone, one column per warp iteration with serialised L2 loads: measured with the in-kernel
// timeline (profiles/ktime_probe.py) that tail cost 28 us at j = 128 (15 us at j = 64) in k_multidot AND in
// k_axpy_dot -- ~5 % of a step at the per-GPU share of N = 8.  Here the fold is a two-level tree that runs while
// the other CTAs are still streaming:
//   level A  CTAs are grouped by blockIdx in groups of RT_GROUP; the last CTA of a group to arrive (group ticket)
//            sums the group's rows -- one column per thread, RT_GROUP independent coalesced loads -- into a group row;
//   level B  the last group to finish (global ticket) sums the <= 64 group rows the same way into `out`.
// The order of every sum is a fixed function of (grid size, jp): bitwise run-to-run deterministic, as before.
// Only two short dependent L2 round trips remain after the last CTA's own stage 1 (~2-3 us).
#pragma once
#include "lkb_kernels.h"

namespace lkb {

enum { RT_GROUP = 16, RT_MAXROWS = 960 };      // RT_MAXROWS + RT_MAXROWS / RT_GROUP <= MAX_ROWBLOCKS rows of `partial`

LKB_DI double  ldcg_w(const double* p)  { return __ldcg(p); }
LKB_DI double2 ldcg_w(const double2* p) { return __ldcg(p); }

// Called by ALL threads of EVERY CTA after the CTA stored its row partial[blockIdx.x][0..jp).  counter: 1 + #groups
// unsigned words, zero on entry, zero again on exit.  Returns true (for all threads) in the one CTA that wrote out[].
template <typename W>
LKB_DI bool reduce_rows_tree(W* __restrict__ partial, int jp, W* __restrict__ out, unsigned* __restrict__ counter)
{
    __shared__ bool s_last;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int nb = gridDim.x;
    const int G = (nb + RT_GROUP - 1) / RT_GROUP;
    const int g = blockIdx.x / RT_GROUP;
    const int r0 = g * RT_GROUP;
    const int gsize = min((int)RT_GROUP, nb - r0);
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(&counter[1 + g], 1u) == (unsigned)(gsize - 1));
    __syncthreads();
    if (!s_last) return false;
    __threadfence();
    W* grow = partial + (size_t)(nb + g) * jp;
    for (int col = tid; col < jp; col += nt) {
        W v[RT_GROUP];
#pragma unroll
        for (int r = 0; r < RT_GROUP; ++r)
            v[r] = (r < gsize) ? ldcg_w(&partial[(size_t)(r0 + r) * jp + col]) : zero_v(W());
        W a = v[0];
#pragma unroll
        for (int r = 1; r < RT_GROUP; ++r) wadd(a, v[r]);
        grow[col] = a;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(&counter[0], 1u) == (unsigned)(G - 1));
    __syncthreads();
    if (!s_last) return false;
    __threadfence();
    const W* grows = partial + (size_t)nb * jp;
    for (int col = tid; col < jp; col += nt) {
        W a = zero_v(W());
        for (int g0 = 0; g0 < G; g0 += 8) {
            W v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = (g0 + u < G) ? ldcg_w(&grows[(size_t)(g0 + u) * jp + col]) : zero_v(W());
#pragma unroll
            for (int u = 0; u < 8; ++u) wadd(a, v[u]);
        }
        out[col] = a;
    }
    for (int i = tid; i < G + 1; i += nt) counter[i] = 0u;
    __syncthreads();
    return true;
}

// One double per CTA (norms): all threads of the LAST CTA call this after the ticket; the total is returned in
// thread 0 (fixed order: strided per-thread sums, warp tree, then the warps in order).
LKB_DI double reduce_scalar_last(const double* __restrict__ partial, int nb)
{
    __shared__ double s_red[32];
    const int tid = threadIdx.x, nt = blockDim.x;
    double a = 0.0;
    for (int b = tid; b < nb; b += nt) a += __ldcg(&partial[b]);
    a = warp_sum(a);
    __syncthreads();
    if ((tid & 31) == 0) s_red[tid >> 5] = a;
    __syncthreads();
    double t = 0.0;
    if (tid == 0)
        for (int q = 0; q < (nt + 31) / 32; ++q) t += s_red[q];
    return t;
}

}  // namespace lkb
