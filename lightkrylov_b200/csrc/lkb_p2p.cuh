// lkb_p2p.cuh -- allreduce fused into the second reduction stage of the multi-dot / multi-axpy /
// fused kernels: the last CTA of each rank pushes its (j+1) partial coefficients straight into
// every peer's HBM over NVLink (P2P stores), raises a per-rank epoch flag, waits for the peers'
// flags and sums the world slots in rank order.  Latency-bound messages (<= 16 KB): this replaces
// a separate ncclAllReduce launch (~20 us) by a few NVLink round trips inside the same kernel.
// Result is bitwise identical on every rank (fixed rank order) and run-to-run deterministic.
// Two data buffers alternate by epoch parity: a rank can run at most one collective ahead of the
// slowest peer, because completing epoch e+1 requires every peer's contribution to e+1.
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

// Called by ALL threads of one CTA.  vals[0..count) (global, W type) holds this rank's sums on
// entry and the world total on exit.  count <= P2P_SLOT.
template <typename W>
LKB_DI void p2p_allreduce_chunk(const P2P& c, W* vals, int count);

// count may exceed the slot size (j + 1 > P2P_SLOT coefficients): processed in slot-sized rounds, one
// epoch each, identically on every rank.
template <typename W>
LKB_DI void p2p_allreduce_cta(const P2P& c, W* vals, int count) {
    for (int base = 0; base < count; base += P2P_SLOT)
        p2p_allreduce_chunk<W>(c, vals + base, min((int)P2P_SLOT, count - base));
}

template <typename W>
LKB_DI void p2p_allreduce_chunk(const P2P& c, W* vals, int count) {
    __shared__ unsigned s_epoch;
    const int tid = threadIdx.x, nth = blockDim.x;
    if (tid == 0) s_epoch = *c.epoch + 1u;
    __syncthreads();
    const unsigned ep = s_epoch;
    const int buf = (int)(ep & 1u);
    const size_t slot_bytes = (size_t)P2P_SLOT * 16;
    // 1. push my values into slot [buf][my rank] of every rank (16-byte words; real kinds use .x)
    for (int i = tid; i < count; i += nth) {
        double2 v = make_double2(0.0, 0.0);
        if constexpr (sizeof(W) == 16) v = *reinterpret_cast<const double2*>(&vals[i]);
        else v.x = *reinterpret_cast<const double*>(&vals[i]);
        for (int r = 0; r < c.world; ++r) {
            double2* dst = reinterpret_cast<double2*>(c.peer[r] + P2P_FLAG_BYTES + ((size_t)buf * P2P_MAXW + c.rank) * slot_bytes);
            dst[i] = v;
        }
    }
    __threadfence_system();
    __syncthreads();
    // 2. raise my flag in every rank, 3. wait for every rank's flag in my region
    if (tid < c.world) st_volatile_u32(reinterpret_cast<unsigned*>(c.peer[tid] + (size_t)c.rank * 128), ep);
    if (tid < c.world) {
        spin_until(reinterpret_cast<const unsigned*>(c.peer[c.rank] + (size_t)tid * 128), ep);
    }
    __syncthreads();
    __threadfence_system();
    // 4. fixed rank-order sum of the world slots
    for (int i = tid; i < count; i += nth) {
        double2 a = make_double2(0.0, 0.0);
        for (int r = 0; r < c.world; ++r) {
            const double2* src = reinterpret_cast<const double2*>(c.peer[c.rank] + P2P_FLAG_BYTES + ((size_t)buf * P2P_MAXW + r) * slot_bytes);
            const double2 v = ld_cv_w(src + i);
            a.x += v.x; a.y += v.y;
        }
        if constexpr (sizeof(W) == 16) *reinterpret_cast<double2*>(&vals[i]) = a;
        else *reinterpret_cast<double*>(&vals[i]) = a.x;
    }
    if (tid == 0) *c.epoch = ep;
    __syncthreads();                     // s_epoch may be rewritten by the next round
}

}  // namespace lkb
