// lkb_p2p.cuh -- allreduce fused into the second reduction stage of the multi-dot / multi-axpy /
// fused kernels: the last CTA of each rank pushes its (j+1) partial coefficients straight into
// every peer's HBM over NVLink (P2P stores), raises a per-rank epoch flag, waits for the peers'
// flags and sums the world slots in rank order.  Latency-bound messages (<= 16 KB): this replaces
// a separate ncclAllReduce launch (~20 us) by a few NVLink round trips inside the same kernel.
// Result is bitwise identical on every rank (fixed rank order) and run-to-run deterministic.
// Two data buffers alternate by epoch parity: a rank can run at most one collective ahead of the
// slowest peer, because completing epoch e+1 requires every peer's contribution to e+1 (which that peer sends
// only after it has finished reading epoch e).
#pragma once
#include "lkb_kernels.h"

namespace lkb {

LKB_DI unsigned long long gtimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// record a timeline point of this CTA (thread 0 only; no-op unless lkb_debug_ktime enabled the buffer)
LKB_DI void ktime_cta(const P2P& c, int slot) {
    if (c.dbg && threadIdx.x == 0) c.dbg[(size_t)blockIdx.x * 4 + slot] = gtimer_ns();
}
LKB_DI void ktime_last(const P2P& c, int slot) {
    if (c.dbg && threadIdx.x == 0) c.dbg[(size_t)4 * MAX_ROWBLOCKS + slot] = gtimer_ns();
}
// running totals over all launches since lkb_debug_ktime(enable): words [4096 + 8 + 2*cls] += allreduce time (ns),
// [.. + 1] += 1, cls 0 = k_multidot, 1 = k_axpy_dot (how long the last CTA spent in the cross-GPU exchange, i.e.
// NVLink latency + waiting for the slowest rank)
LKB_DI void ktime_accumulate_wait(const P2P& c, int cls) {
    if (c.dbg && threadIdx.x == 0) {
        unsigned long long* t = c.dbg + (size_t)4 * MAX_ROWBLOCKS;
        t[8 + 2 * cls] += t[2] - t[1];
        t[9 + 2 * cls] += 1ULL;
    }
}
LKB_DI unsigned ld_volatile_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
LKB_DI void st_volatile_u32(unsigned* p, unsigned v) {
    asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// Spin until *f reaches epoch ep.  A peer that died (or a collective entered by only some ranks) must not
// hang the GPU: after ~30 s of spinning the kernel traps, which surfaces as a CUDA launch failure on the host.
LKB_DI void spin_until(const unsigned* f, unsigned ep) {
    const long long t0 = clock64();
    while ((int)(ld_volatile_u32(f) - ep) < 0) {
        if (clock64() - t0 > 60000000000LL) { asm volatile("trap;"); }
    }
}
LKB_DI double2 ld_cv_w(const double2* p) {
    double2 v;
    asm volatile("ld.volatile.global.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p) : "memory");
    return v;
}

// p2p_allreduce_cta: called by ALL threads of one CTA.  vals[0..count) (global, W type) holds this rank's sums on entry
// and the world total on exit (bitwise identical on every rank: fixed rank order).
// Low-latency protocol (round 2): every 16-byte word carries its own arrival tags, so a receiver spins on the words
// themselves.  No __threadfence_system + separate flag store: one one-way NVLink latency per collective instead of a
// fence round trip plus a flag hop.  This is NCCL's "LL" layout: each 8-byte half of the word is {32 bits of the
// double, 32-bit epoch tag} and is valid on its own, so nothing depends on the two halves of the 128-bit store
// becoming visible together (8-byte accesses are single-copy atomic).  Complex coefficients take two words.
// Slots alternate by epoch parity; a stale word of the same parity carries tag epoch - 2.
LKB_DI void st_ll(double2* p, double v, unsigned ep) {
    const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
    const unsigned long long tag = (unsigned long long)ep << 32;
    const unsigned long long lo = (bits & 0xffffffffULL) | tag, hi = (bits >> 32) | tag;
    asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(lo), "l"(hi) : "memory");
}
LKB_DI double ld_ll(const double2* p, unsigned ep) {
    unsigned long long lo, hi;
    const long long t0 = clock64();
    while (true) {
        asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(lo), "=l"(hi) : "l"(p) : "memory");
        if ((unsigned)(lo >> 32) == ep && (unsigned)(hi >> 32) == ep) break;
        if (clock64() - t0 > 60000000000LL) { asm volatile("trap;"); }       // a peer died: fail instead of hanging the GPU
    }
    return __longlong_as_double((long long)((lo & 0xffffffffULL) | (hi << 32)));
}

template <typename W>
LKB_DI void p2p_allreduce_chunk(const P2P& c, W* vals, int count) {
    __shared__ unsigned s_epoch;
    constexpr int WPV = sizeof(W) / 8;                  // 16-byte words per value: 1 (real) or 2 (re, im)
    const int tid = threadIdx.x, nth = blockDim.x;
    if (tid == 0) s_epoch = *c.epoch + 1u;
    __syncthreads();
    const unsigned ep = s_epoch;
    const int buf = (int)(ep & 1u);
    const size_t slot_bytes = (size_t)P2P_SLOT * 16;
    const int nwords = count * WPV;
    const double* src = reinterpret_cast<const double*>(vals);
    // 1. push {value, epoch} words into slot [buf][my rank] of every rank
    for (int i = tid; i < nwords; i += nth) {
        const double v = src[i];
        for (int r = 0; r < c.world; ++r) {
            double2* dst = reinterpret_cast<double2*>(c.peer[r] + P2P_FLAG_BYTES + ((size_t)buf * P2P_MAXW + c.rank) * slot_bytes);
            st_ll(dst + i, v, ep);
        }
    }
    // 2. every word of every rank's slot in MY region: spin until its tag is this epoch, sum in rank order
    double* out = reinterpret_cast<double*>(vals);
    for (int i = tid; i < nwords; i += nth) {
        double a = 0.0;
        for (int r = 0; r < c.world; ++r) {
            const double2* s = reinterpret_cast<const double2*>(c.peer[c.rank] + P2P_FLAG_BYTES + ((size_t)buf * P2P_MAXW + r) * slot_bytes);
            a += ld_ll(s + i, ep);
        }
        out[i] = a;
    }
    __syncthreads();
    if (tid == 0) *c.epoch = ep;
    __syncthreads();                     // s_epoch may be rewritten by the next round
}

// count may exceed the slot size: processed in slot-sized rounds, one epoch each, identically on every rank.
template <typename W>
LKB_DI void p2p_allreduce_cta(const P2P& c, W* vals, int count) {
    constexpr int PER = P2P_SLOT / (int)(sizeof(W) / 8);
    for (int base = 0; base < count; base += PER)
        p2p_allreduce_chunk<W>(c, vals + base, min(PER, count - base));
}

}  // namespace lkb
