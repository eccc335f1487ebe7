// kernels_ops.cu -- abstract_linop extensions: matrix-free 5-/7-point stencils and CSR SpMV.
//
// Reference contract: src/AbstractTypes/AbstractLinops.fypp:58-87 (matvec / rmatvec, vec_out is
// intent(out)).  Stencil index convention: p = i + nx*(j + ny*k); coef = (center,-x,+x,-y,+y,-z,+z),
// homogeneous Dirichlet outside the global grid.  The slowest axis (y in 2-D, z in 3-D) is the
// row-sharded one: the neighbouring slab's boundary row / plane arrives in halo_lo / halo_hi.
#include <stdlib.h>
#include <algorithm>
#include "lkb_kernels.h"
#include "lkb_p2p.cuh"

namespace lkb {

template <typename E> struct Coef7 { E c[7]; };

template <typename E, int PW> struct PackOps {
    using P = Pack<E, PW>;
    // Pack is alignas(16), so sizeof(P) == 16 even for a partial pack (PW elements < 16 bytes: rows whose length is not a
    // multiple of the pack width): only a FULL pack may be moved with one 128-bit access
    static constexpr bool FULL = (PW * sizeof(E) == 16);
    static LKB_DI P ld(const E* p) {
        if constexpr (FULL) return ld_pack_l1<P>(p);
        else { P r;
#pragma unroll
            for (int e = 0; e < PW; ++e) r.v[e] = __ldg(p + e);
            return r; }
    }
    // halo data written by a peer GPU during this kernel's lifetime: bypass L1 (ld.global.cg)
    static LKB_DI P ld_cg(const E* p) {
        if constexpr (FULL) { int4 r = __ldcg(reinterpret_cast<const int4*>(p)); return *reinterpret_cast<P*>(&r); }
        else { P r;
#pragma unroll
            for (int e = 0; e < PW; ++e) r.v[e] = __ldcg(p + e);
            return r; }
    }
    static LKB_DI void st(E* p, const P& v) {
        if constexpr (FULL) st_pack(p, v);
        else {
#pragma unroll
            for (int e = 0; e < PW; ++e) p[e] = v.v[e];
        }
    }
    // the matvec output is the work vector w of the step that follows: small slices are stored with L2 evict_last (lkb_types.cuh)
    static LKB_DI void st_w(E* p, const P& v, bool keep, uint64_t pol) {
        if constexpr (FULL) st_pack_w(p, v, keep, pol);
        else st(p, v);
    }
    static LKB_DI P zero() { P r;
#pragma unroll
        for (int e = 0; e < PW; ++e) r.v[e] = zero_v(E());
        return r; }
};


// Block decode shared by the two stencil kernels.  1-D grid, column block fastest (consecutive CTAs stream
// consecutive segments of the same grid rows); the row groups are rotated by one so that the two groups that
// read the inter-GPU halos (first / last row group in 2-D, first / last plane in 3-D) are scheduled LAST: the
// slab interior is computed while the neighbours' halo rows are still in flight ("overlapped with the
// interior update").  halo_wait: the halo-reading CTAs wait until the neighbours published the epoch the local
// counter says (k_halo_push has already waited, so this is one load; after a push fused into k_multiaxpy_fin /
// k_scale_dev nobody has waited yet and this is where the exchange is synchronised).
struct StBlock { int64_t cb, rb, k; };
template <int DIM>
LKB_DI StBlock st_decode(int64_t ncb, int64_t nyb, int64_t nz) {
    StBlock b;
    const int64_t bid = blockIdx.x;
    b.cb = bid % ncb;
    const int64_t rg = bid / ncb;
    if (DIM == 2) { b.k = 0; b.rb = (rg + 1) % nyb; }
    else { b.rb = rg % nyb; b.k = (rg / nyb + 1) % nz; }
    return b;
}
LKB_DI void halo_wait(const unsigned* halo_epoch, const unsigned* flag_lo, const unsigned* flag_hi, bool need_lo, bool need_hi) {
    if (!halo_epoch || !(need_lo || need_hi)) return;      // uniform over the CTA
    if (threadIdx.x == 0) {
        const unsigned ep = *halo_epoch;
        if (need_lo && flag_lo) spin_until(flag_lo, ep);
        if (need_hi && flag_hi) spin_until(flag_hi, ep);
        __threadfence_system();
    }
    __syncthreads();
}

// One CTA = 256 threads side by side along x (256*PW points of one grid row), marching RY rows in y
// (2.5-D blocking): the south / centre / north packs of a column live in registers, so every x
// element is loaded from HBM once per CTA (+2 halo rows per RY rows, served by L2 because the
// neighbouring CTA streams them at the same time).  The x-halo (west / east neighbour of a pack)
// comes from L1 (the neighbouring thread loaded that line) or, with SHFL, through warp shuffles.
// UN > 1 prefetches UN north rows up front.  Inter-GPU halos arrive in halo_lo / halo_hi (slab
// edges, filled by ncclSend/Recv).  HBM traffic ~ (2 + 2/RY) * n * s.
template <int K, int PW, int DIM, int RY, int UN, bool SHFL>
__global__ void __launch_bounds__(256)
k_stencil(const typename Tr<K>::E* __restrict__ x, typename Tr<K>::E* __restrict__ y,
          int64_t nx, int64_t ny, int64_t nz, Coef7<typename Tr<K>::E> cf,
          const typename Tr<K>::E* __restrict__ halo_lo, const typename Tr<K>::E* __restrict__ halo_hi,
          const unsigned* __restrict__ halo_epoch, int64_t halo_parity_stride, const unsigned* __restrict__ flag_lo,
          const unsigned* __restrict__ flag_hi, int64_t ncb, const int* __restrict__ flags, const int keep_w)
{
    using E = typename Tr<K>::E;
    using PO = PackOps<E, PW>;
    const bool keep = keep_w != 0;
    const uint64_t pol = keep ? pol_evict_last() : 0ULL;
    pdl_wait();            // x, the stop flag and the halo epoch come from the predecessor
    pdl_trigger();
    if (flags && flags[F_STOP]) return;
    if (halo_epoch) {      // double-buffered p2p halos: pick the parity of the current epoch
        const int64_t off = (int64_t)(*halo_epoch & 1u) * halo_parity_stride;
        if (halo_lo) halo_lo += off;
        if (halo_hi) halo_hi += off;
    }
    using P = typename PO::P;
    const int64_t npk_row = nx / PW;
    const int64_t nyb = (ny + RY - 1) / RY;
    const StBlock sb = st_decode<DIM>(ncb, nyb, nz);
    const int64_t ip = sb.cb * blockDim.x + threadIdx.x;
    const bool active = ip < npk_row;                 // inactive lanes still take part in the shuffles
    const int64_t i0 = (active ? ip : npk_row - 1) * PW;
    const int64_t k = sb.k;
    const int64_t j0 = sb.rb * RY;
    const int64_t j1 = min(ny, j0 + RY);
    halo_wait(halo_epoch, flag_lo, flag_hi, halo_lo && (DIM == 2 ? j0 == 0 : k == 0),
              halo_hi && (DIM == 2 ? j1 >= ny : k == nz - 1));
    const int64_t plane = nx * ny;
    const E* xk = x + k * plane;
    const int lane = threadIdx.x & 31;

    auto getrow = [&](int64_t j) -> P {
        if (DIM == 2) {
            if (j < 0) return halo_lo ? PO::ld_cg(halo_lo + i0) : PO::zero();
            if (j >= ny) return halo_hi ? PO::ld_cg(halo_hi + i0) : PO::zero();
        } else {
            if (j < 0 || j >= ny) return PO::zero();
        }
        return PO::ld(xk + j * nx + i0);
    };
    auto shfl_e = [&](E v, int src) -> E {
        if constexpr (sizeof(E) == 4) return __shfl_sync(0xffffffffu, v, src);
        else if constexpr (sizeof(E) == 8 && !Tr<K>::cplx) return __shfl_sync(0xffffffffu, v, src);
        else if constexpr (sizeof(E) == 8) { E r; r.x = __shfl_sync(0xffffffffu, v.x, src); r.y = __shfl_sync(0xffffffffu, v.y, src); return r; }
        else { E r; r.x = __shfl_sync(0xffffffffu, v.x, src); r.y = __shfl_sync(0xffffffffu, v.y, src); return r; }
    };
    auto do_row = [&](int64_t j, const P& south, const P& center, const P& north) {
        const int64_t p = j * nx + i0;
        // x-halo through warp shuffles; edge lanes of the warp (and of the row) load / zero-fill
        E west, east;
        if (SHFL) {
            west = shfl_e(center.v[PW - 1], (lane + 31) & 31);
            east = shfl_e(center.v[0], (lane + 1) & 31);
            if (lane == 0) west = i0 > 0 ? __ldg(xk + p - 1) : zero_v(E());
            if (lane == 31 || ip + 1 >= npk_row) east = (i0 + PW < nx) ? __ldg(xk + p + PW) : zero_v(E());
        } else {
            west = i0 > 0 ? __ldg(xk + p - 1) : zero_v(E());
            east = (i0 + PW < nx) ? __ldg(xk + p + PW) : zero_v(E());
        }
        P down, up;
        if (DIM == 3) {
            down = (k > 0) ? PO::ld(xk + p - plane) : (halo_lo ? PO::ld_cg(halo_lo + p) : PO::zero());
            up = (k < nz - 1) ? PO::ld(xk + p + plane) : (halo_hi ? PO::ld_cg(halo_hi + p) : PO::zero());
        }
        P out;
#pragma unroll
        for (int e = 0; e < PW; ++e) {
            E s = mul_v(cf.c[0], center.v[e]);
            fmacc(s, (e > 0 ? center.v[e > 0 ? e - 1 : 0] : west), cf.c[1]);
            fmacc(s, (e < PW - 1 ? center.v[e < PW - 1 ? e + 1 : 0] : east), cf.c[2]);
            fmacc(s, south.v[e], cf.c[3]);
            fmacc(s, north.v[e], cf.c[4]);
            if (DIM == 3) { fmacc(s, down.v[e], cf.c[5]); fmacc(s, up.v[e], cf.c[6]); }
            out.v[e] = s;
        }
        if (active) PO::st_w(y + k * plane + p, out, keep, pol);
    };

    P south = getrow(j0 - 1), center = getrow(j0);
    int64_t j = j0;
    for (; j + UN <= j1; j += UN) {
        P nb[UN];
#pragma unroll
        for (int u = 0; u < UN; ++u) nb[u] = getrow(j + 1 + u);
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            do_row(j + u, south, center, nb[u]);
            south = center; center = nb[u];
        }
    }
    for (; j < j1; ++j) {
        const P north = getrow(j + 1);
        do_row(j, south, center, north);
        south = center; center = north;
    }
}

// ------------------------------------------------------------------------------------------
// Shared-memory halo-staged stencil (default).  One CTA = ST_TX packs in x by RY rows in y.  The CTA
// first issues ALL its loads -- the (RY+2) x ST_TX tile of plane k including the y-halo rows (from
// the neighbouring rows, from the inter-GPU halo buffers at slab edges, or zero at the Dirichlet
// boundary) and, in 3-D, the RY x ST_TX centre rows of planes k-1 / k+1 -- as 16-byte cp.async
// copies straight into shared memory (RY+2 .. 3RY+2 requests in flight per thread, no register
// staging), waits once, and then computes the RY output rows from shared memory: north / south /
// west / east neighbours are shared-memory reads, only the west/east scalar of the tile's edge
// columns comes from global memory.  Every x element is fetched from L2/HBM once per CTA.
// ------------------------------------------------------------------------------------------
enum { ST_TX = 128 };

template <int BYTES> LKB_DI void cp_async(void* smem_dst, const void* gsrc) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    if constexpr (BYTES == 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
    else if constexpr (BYTES == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
    else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gsrc) : "memory");
}
LKB_DI void cp_async_wait_all() { asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory"); }

template <int K, int PW, int DIM, int RY>
__global__ void __launch_bounds__(ST_TX)
k_stencil_smem(const typename Tr<K>::E* __restrict__ x, typename Tr<K>::E* __restrict__ y,
               int64_t nx, int64_t ny, int64_t nz, Coef7<typename Tr<K>::E> cf,
               const typename Tr<K>::E* __restrict__ halo_lo, const typename Tr<K>::E* __restrict__ halo_hi,
               const unsigned* __restrict__ halo_epoch, int64_t halo_parity_stride, const unsigned* __restrict__ flag_lo,
               const unsigned* __restrict__ flag_hi, int64_t ncb, const int* __restrict__ flags, const int keep_w)
{
    using E = typename Tr<K>::E;
    using P = Pack<E, PW>;
    constexpr int PB = PW * (int)sizeof(E);       // payload bytes (Pack is padded to 16 B)
    const bool keep = keep_w != 0;
    const uint64_t pol = keep ? pol_evict_last() : 0ULL;
    pdl_wait();
    pdl_trigger();
    if (flags && flags[F_STOP]) return;
    if (halo_epoch) {
        const int64_t off = (int64_t)(*halo_epoch & 1u) * halo_parity_stride;
        if (halo_lo) halo_lo += off;
        if (halo_hi) halo_hi += off;
    }
    extern __shared__ __align__(16) unsigned char st_smem[];
    P* tile = reinterpret_cast<P*>(st_smem);                       // [(RY+2)][ST_TX]
    P* dn = tile + (RY + 2) * ST_TX;                               // [RY][ST_TX]  (3-D only)
    P* up = dn + RY * ST_TX;                                       // [RY][ST_TX]
    const int tx = threadIdx.x;
    const int64_t npk_row = nx / PW;
    const int64_t nyb = (ny + RY - 1) / RY;
    const StBlock sb = st_decode<DIM>(ncb, nyb, nz);
    const int64_t ip = sb.cb * ST_TX + tx;
    const bool active = ip < npk_row;
    const int64_t i0 = ip * PW;
    const int64_t k = sb.k;
    const int64_t j0 = sb.rb * RY;
    const int64_t plane = nx * ny;
    halo_wait(halo_epoch, flag_lo, flag_hi, halo_lo && (DIM == 2 ? j0 == 0 : k == 0),
              halo_hi && (DIM == 2 ? j0 + RY >= ny : k == nz - 1));
    const E* xk = x + k * plane;
    P zero;
#pragma unroll
    for (int e = 0; e < PW; ++e) zero.v[e] = zero_v(E());

    if (active) {
        // ---- issue every load of the tile ----
#pragma unroll
        for (int r = 0; r < RY + 2; ++r) {
            const int64_t j = j0 - 1 + r;
            const E* src = nullptr;
            if (j >= 0 && j < ny) src = xk + j * nx + i0;
            else if (DIM == 2 && j < 0 && halo_lo) src = halo_lo + i0;
            else if (DIM == 2 && j >= ny && j == ny && halo_hi) src = halo_hi + i0;
            if (src && j <= ny) cp_async<PB>(&tile[r * ST_TX + tx], src);
            else tile[r * ST_TX + tx] = zero;
        }
        if (DIM == 3) {
#pragma unroll
            for (int r = 0; r < RY; ++r) {
                const int64_t j = j0 + r;
                if (j < ny) {
                    const int64_t p = j * nx + i0;
                    const E* sd = (k > 0) ? xk + p - plane : (halo_lo ? halo_lo + p : nullptr);
                    const E* su = (k < nz - 1) ? xk + p + plane : (halo_hi ? halo_hi + p : nullptr);
                    if (sd) cp_async<PB>(&dn[r * ST_TX + tx], sd); else dn[r * ST_TX + tx] = zero;
                    if (su) cp_async<PB>(&up[r * ST_TX + tx], su); else up[r * ST_TX + tx] = zero;
                }
            }
        }
    }
    cp_async_wait_all();
    __syncthreads();
    if (!active) return;
    // ---- compute RY rows out of shared memory ----
#pragma unroll
    for (int r = 0; r < RY; ++r) {
        const int64_t j = j0 + r;
        if (j >= ny) break;
        const int64_t p = j * nx + i0;
        const P center = tile[(r + 1) * ST_TX + tx], south = tile[r * ST_TX + tx], north = tile[(r + 2) * ST_TX + tx];
        E west, east;
        if (tx > 0) west = tile[(r + 1) * ST_TX + tx - 1].v[PW - 1];
        else west = i0 > 0 ? __ldg(xk + p - 1) : zero_v(E());
        if (tx < ST_TX - 1 && ip + 1 < npk_row) east = tile[(r + 1) * ST_TX + tx + 1].v[0];
        else east = (i0 + PW < nx) ? __ldg(xk + p + PW) : zero_v(E());
        P out;
#pragma unroll
        for (int e = 0; e < PW; ++e) {
            E sacc = mul_v(cf.c[0], center.v[e]);
            fmacc(sacc, (e > 0 ? center.v[e > 0 ? e - 1 : 0] : west), cf.c[1]);
            fmacc(sacc, (e < PW - 1 ? center.v[e < PW - 1 ? e + 1 : 0] : east), cf.c[2]);
            fmacc(sacc, south.v[e], cf.c[3]);
            fmacc(sacc, north.v[e], cf.c[4]);
            if (DIM == 3) {
                fmacc(sacc, dn[r * ST_TX + tx].v[e], cf.c[5]);
                fmacc(sacc, up[r * ST_TX + tx].v[e], cf.c[6]);
            }
            out.v[e] = sacc;
        }
        PackOps<E, PW>::st_w(y + k * plane + p, out, keep, pol);
    }
}


// ------------------------------------------------------------------------------------------
// 3-D z-march (round 2; "3.5-D blocking").  One CTA = ZT_TX packs in x by ZT_TY rows in y, marching LZ planes in z
// with the below / centre / above packs of its column in REGISTERS and the next plane prefetched one iteration
// ahead.  The y-neighbours (rows j-1 / j+1 of the centre plane) are re-read through L1: the neighbouring threads
// of the same CTA loaded exactly those lines one iteration earlier as their "above" pack.  Per point and plane the
// kernel issues ONE load that can miss L1 -- against three in the y-marching kernel, whose z-neighbours belong to
// other CTAs and are always served by L2 (ncu, 384^3 fp64: DRAM traffic already equal to the algorithmic 2 n s, but
// 3.4 n s of L2->L1 traffic and 4.6 TB/s).  No shared memory, no block barrier (the shared-memory z-march of round 1
// paid two barriers per plane: 2.8 TB/s).  Inter-GPU halo planes (k = -1, k = nz) come from halo_lo / halo_hi.
// ------------------------------------------------------------------------------------------
enum { ZT_TX = 32, ZT_TY = 8 };

template <int K, int PW, int LZ>
__global__ void __launch_bounds__(ZT_TX * ZT_TY)
k_stencil3d_zmarch(const typename Tr<K>::E* __restrict__ x, typename Tr<K>::E* __restrict__ y,
                   int64_t nx, int64_t ny, int64_t nz, Coef7<typename Tr<K>::E> cf,
                   const typename Tr<K>::E* __restrict__ halo_lo, const typename Tr<K>::E* __restrict__ halo_hi,
                   const unsigned* __restrict__ halo_epoch, int64_t halo_parity_stride, const unsigned* __restrict__ flag_lo,
                   const unsigned* __restrict__ flag_hi, int64_t ncbx, int64_t ncby, const int* __restrict__ flags)
{
    using E = typename Tr<K>::E;
    using PO = PackOps<E, PW>;
    using P = typename PO::P;
    pdl_wait();
    pdl_trigger();
    if (flags && flags[F_STOP]) return;
    if (halo_epoch) {
        const int64_t off = (int64_t)(*halo_epoch & 1u) * halo_parity_stride;
        if (halo_lo) halo_lo += off;
        if (halo_hi) halo_hi += off;
    }
    const int64_t nzc = (nz + LZ - 1) / LZ;
    const int64_t bid = blockIdx.x;
    const int64_t cbx = bid % ncbx, cby = (bid / ncbx) % ncby;
    const int64_t zc = ((bid / (ncbx * ncby)) + 1) % nzc;            // the chunks touching the inter-GPU halos run last
    const int64_t k0 = zc * LZ, k1 = min(nz, k0 + LZ);
    halo_wait(halo_epoch, flag_lo, flag_hi, halo_lo && k0 == 0, halo_hi && k1 == nz);
    const int tx = threadIdx.x % ZT_TX, ty = threadIdx.x / ZT_TX;
    const int64_t npk_row = nx / PW;
    const int64_t ip = cbx * ZT_TX + tx;
    const int64_t j = cby * ZT_TY + ty;
    if (ip >= npk_row || j >= ny) return;
    const int64_t i0 = ip * PW;
    const int64_t plane = nx * ny;
    const int64_t q = j * nx + i0;                                   // offset inside a plane
    auto load_plane = [&](int64_t k) -> P {
        if (k < 0) return halo_lo ? PO::ld_cg(halo_lo + q) : PO::zero();
        if (k >= nz) return halo_hi ? PO::ld_cg(halo_hi + q) : PO::zero();
        return PO::ld(x + k * plane + q);
    };
    P down = load_plane(k0 - 1), center = load_plane(k0), up = load_plane(k0 + 1);
    for (int64_t k = k0; k < k1; ++k) {
        const P nxt = (k + 2 <= k1) ? load_plane(k + 2) : PO::zero();          // prefetch: used two iterations later as `up`
        const E* xc = x + k * plane + q;
        const P north = (j + 1 < ny) ? PO::ld(xc + nx) : PO::zero();
        const P south = (j > 0) ? PO::ld(xc - nx) : PO::zero();
        const E west = i0 > 0 ? __ldg(xc - 1) : zero_v(E());
        const E east = (i0 + PW < nx) ? __ldg(xc + PW) : zero_v(E());
        P out;
#pragma unroll
        for (int e = 0; e < PW; ++e) {
            E s = mul_v(cf.c[0], center.v[e]);
            fmacc(s, (e > 0 ? center.v[e > 0 ? e - 1 : 0] : west), cf.c[1]);
            fmacc(s, (e < PW - 1 ? center.v[e < PW - 1 ? e + 1 : 0] : east), cf.c[2]);
            fmacc(s, south.v[e], cf.c[3]);
            fmacc(s, north.v[e], cf.c[4]);
            fmacc(s, down.v[e], cf.c[5]);
            fmacc(s, up.v[e], cf.c[6]);
            out.v[e] = s;
        }
        PO::st(y + k * plane + q, out);
        down = center; center = up; up = nxt;
    }
}

template <int K, int PW, int DIM>
static void stencil_launch(cudaStream_t s, const StencilArgs& a, const void* x, void* y, bool trans, const int* flags) {
    using E = typename Tr<K>::E;
    Coef7<E> cf;
    Scalar c[7];
    for (int q = 0; q < 7; ++q) c[q] = a.coef[q];
    if (trans) {  // A^H of a constant-coefficient stencil: swap +/- neighbours, conjugate
        Scalar t;
        t = c[1]; c[1] = c[2]; c[2] = t; t = c[3]; c[3] = c[4]; c[4] = t; t = c[5]; c[5] = c[6]; c[6] = t;
        for (int q = 0; q < 7; ++q) c[q].im = -c[q].im;
    }
    for (int q = 0; q < 7; ++q) from_scalar(c[q], cf.c[q]);
    const int64_t npk_row = a.nx / PW;
    const int keep_w = ((size_t)a.nx * (size_t)a.ny * (size_t)(DIM == 3 ? a.nz : 1) * sizeof(E) <= w_keep_bytes()) ? 1 : 0;
    // Measured on B200 (profiles/stencil_ab.py, fp64): 2-D 4096^2 -- shared-memory kernel RY=8 5.64 TB/s vs
    // register march 5.24 TB/s; 3-D 384^3 -- shared-memory kernel 2.1-3.0 TB/s (3RY+2 staged rows per CTA cut
    // the occupancy) vs register march 4.62 TB/s; a z-march with the plane staged in shared memory (two block
    // barriers per plane) reached only 2.8 TB/s and was dropped; a register z-march (variant 4, k_stencil3d_zmarch) 3.95 TB/s.
    // Default: shared-memory staging in 2-D, register y-march in 3-D.
    // 3-D: the z-march (variant 4) measured 3.95 TB/s against 4.32 TB/s for the y-march at 384^3 in three separate runs
    // (gpurun_out/r02_stencil_ab*.txt) -- the y-march stays the default
    static const int env_variant = getenv("LKB_STENCIL_VARIANT") ? atoi(getenv("LKB_STENCIL_VARIANT")) : -1;
    const int variant = env_variant >= 0 ? env_variant : (a.variant >= 0 ? a.variant : (DIM == 2 ? 1 : 0));
    if constexpr (DIM == 3) {
        if (variant == 4) {      // z-march with register planes (A/B only)
            constexpr int LZ = 32;
            const int64_t ncbx = (npk_row + ZT_TX - 1) / ZT_TX, ncby = (a.ny + ZT_TY - 1) / ZT_TY, nzc = (a.nz + LZ - 1) / LZ;
            const unsigned grid = (unsigned)(ncbx * ncby * nzc);
            launch_ex(k_stencil3d_zmarch<K, PW, LZ>, grid, ZT_TX * ZT_TY, 0, s, pdl_take(1), (const E*)x, (E*)y, a.nx, a.ny, a.nz, cf,
                (const E*)a.halo_lo, (const E*)a.halo_hi, a.halo_epoch, a.halo_parity_stride, a.flag_lo, a.flag_hi, ncbx, ncby, flags);
            return;
        }
    }
    if (variant >= 1 && variant <= 3) {
        // shared-memory halo-staged kernel
#define LKB_STS(RY_) { const int64_t nyb_ = (a.ny + RY_ - 1) / RY_; \
            const int64_t ncb_ = (npk_row + ST_TX - 1) / ST_TX; \
            const unsigned grid_ = (unsigned)(nyb_ * (DIM == 3 ? a.nz : 1) * ncb_); \
            const size_t sh_ = (size_t)((RY_ + 2) + (DIM == 3 ? 2 * RY_ : 0)) * ST_TX * sizeof(Pack<E, PW>); \
            static const SmemAttrOnce once_((const void*)k_stencil_smem<K, PW, DIM, RY_>, 96 * 1024); once_.ensure(); \
            launch_ex(k_stencil_smem<K, PW, DIM, RY_>, grid_, ST_TX, sh_, s, pdl_take(1), (const E*)x, (E*)y, a.nx, a.ny, a.nz, cf, \
                (const E*)a.halo_lo, (const E*)a.halo_hi, a.halo_epoch, a.halo_parity_stride, a.flag_lo, a.flag_hi, ncb_, flags, keep_w); }
        if (variant == 3) LKB_STS(4) else LKB_STS(8)
#undef LKB_STS
        return;
    }
    // variant 0: register-marching kernel.  RY = 8 rows per CTA, no unrolling, L1-served x-halo
    // (profiles/stencil_ab.py, B200: 5.2 TB/s 2-D 4096^2 / 4.6 TB/s 3-D 384^3; longer marches or 4x-unrolled
    // prefetch were slower, warp-shuffle x-halo made no difference).
    constexpr int RY = 8;
    const int64_t nyb = (a.ny + RY - 1) / RY;
    // CTA width: the narrowest multiple of 32 threads that covers a grid row with the same number of CTAs as 256-wide
    // ones would (384^3: 192 packs per row -> 192 threads, no idle lanes; a 256-wide CTA left 25 % of its threads idle:
    // 4.28 TB/s at 384^3 against 5.17 TB/s at 512^3)
    const int64_t ncb = (npk_row + 255) / 256;
    const unsigned bs = (unsigned)std::min<int64_t>(256, ((npk_row + ncb - 1) / ncb + 31) / 32 * 32);
    const unsigned grid = (unsigned)(nyb * (DIM == 3 ? a.nz : 1) * ncb);
    launch_ex(k_stencil<K, PW, DIM, RY, 1, false>, grid, bs, 0, s, pdl_take(1), (const E*)x, (E*)y, a.nx, a.ny, a.nz, cf,
              (const E*)a.halo_lo, (const E*)a.halo_hi, a.halo_epoch, a.halo_parity_stride, a.flag_lo, a.flag_hi, ncb, flags, keep_w);
}

template <int K>
static void stencil_t(cudaStream_t s, const StencilArgs& a, const void* x, void* y, bool trans, const int* flags) {
    constexpr int EPP = Tr<K>::EPP;
    const bool vec = (a.nx % EPP == 0);
    if (a.dim == 2) {
        if (vec) stencil_launch<K, EPP, 2>(s, a, x, y, trans, flags);
        else stencil_launch<K, 1, 2>(s, a, x, y, trans, flags);
    } else {
        if (vec) stencil_launch<K, EPP, 3>(s, a, x, y, trans, flags);
        else stencil_launch<K, 1, 3>(s, a, x, y, trans, flags);
    }
}
void launch_stencil(int kind, cudaStream_t s, const StencilArgs& a, const void* x, void* y, bool trans,
                    const int* flags, int) {
    switch (kind) {
        case KS: stencil_t<KS>(s, a, x, y, trans, flags); break;
        case KD: stencil_t<KD>(s, a, x, y, trans, flags); break;
        case KC: stencil_t<KC>(s, a, x, y, trans, flags); break;
        default: stencil_t<KZ>(s, a, x, y, trans, flags); break;
    }
}

// ------------------------------------------------------------------------------------------
// Halo exchange over NVLink peer memory, fused with its synchronisation: replaces the
// ncclSend/ncclRecv group in front of the stencil.  Every CTA copies a slice of the boundary rows
// straight into the neighbours' (CUDA-IPC mapped) halo buffers; the last CTA to finish raises this
// rank's epoch flag in both neighbours and waits for theirs, so when the kernel retires the local
// halo buffers of the current parity are complete.  Buffers alternate by epoch parity: a neighbour
// can push for matvec m+1 only after it saw my push for matvec m, i.e. after I finished matvec m-1,
// the last reader of that parity.
template <int K>
__global__ void __launch_bounds__(256)
k_halo_push(HaloP2P h, const typename Tr<K>::E* __restrict__ x, int64_t n_loc, const int* __restrict__ flags)
{
    using E = typename Tr<K>::E;
    if (flags && flags[F_STOP]) return;
    __shared__ bool is_last;
    const unsigned ep = *h.epoch + 1u;
    const size_t par = (size_t)(ep & 1u) * 2 * h.side_bytes;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (h.lo_region) {      // my first row/plane -> lower neighbour's halo_hi (side 1)
        E* dst = reinterpret_cast<E*>(h.lo_region + h.data_off + par + h.side_bytes);
        for (int64_t i = t0; i < h.he; i += stride) dst[i] = x[i];
    }
    if (h.hi_region) {      // my last row/plane -> upper neighbour's halo_lo (side 0)
        E* dst = reinterpret_cast<E*>(h.hi_region + h.data_off + par);
        const E* src = x + (n_loc - h.he);
        for (int64_t i = t0; i < h.he; i += stride) dst[i] = src[i];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(h.ticket, 1u) == gridDim.x - 1u);
    __syncthreads();
    if (is_last && threadIdx.x == 0) {
        __threadfence_system();
        if (h.lo_region) st_volatile_u32(reinterpret_cast<unsigned*>(h.lo_region + 128), ep);   // "upper neighbour pushed"
        if (h.hi_region) st_volatile_u32(reinterpret_cast<unsigned*>(h.hi_region), ep);         // "lower neighbour pushed"
        if (h.lo_region) spin_until(reinterpret_cast<const unsigned*>(h.my_region), ep);
        if (h.hi_region) spin_until(reinterpret_cast<const unsigned*>(h.my_region + 128), ep);
        __threadfence_system();
        *h.ticket = 0u;
        *h.epoch = ep;
    }
}
void launch_halo_push(int kind, cudaStream_t s, const HaloP2P& h, const void* x, int64_t n_loc, const int* flags) {
    int nb = (int)((h.he + 2047) / 2048);
    if (nb < 1) nb = 1;
    if (nb > 64) nb = 64;
    switch (kind) {
        case KS: k_halo_push<KS><<<nb, 256, 0, s>>>(h, (const float*)x, n_loc, flags); break;
        case KD: k_halo_push<KD><<<nb, 256, 0, s>>>(h, (const double*)x, n_loc, flags); break;
        case KC: k_halo_push<KC><<<nb, 256, 0, s>>>(h, (const float2*)x, n_loc, flags); break;
        default: k_halo_push<KZ><<<nb, 256, 0, s>>>(h, (const double2*)x, n_loc, flags); break;
    }
}

// ------------------------------------------------------------------------------------------
// CSR SpMV.  LPR lanes cooperate on one row (LPR = 32: warp-per-row; LPR < 32: vector-per-row
// for short rows).  Fixed shuffle tree => deterministic.  rmatvec uses an explicit transposed
// copy built at operator creation (no atomics), with conj_vals for A^H.
// ------------------------------------------------------------------------------------------
template <int K, int LPR>
__global__ void __launch_bounds__(256)
k_csr(int64_t m, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
      const typename Tr<K>::E* __restrict__ val, const typename Tr<K>::E* __restrict__ x,
      typename Tr<K>::E* __restrict__ y, bool conj_vals, const int* __restrict__ flags)
{
    using E = typename Tr<K>::E;
    if (flags && flags[F_STOP]) return;
    const int sub = threadIdx.x % LPR;
    const int64_t rows_per_cta = 256 / LPR;
    for (int64_t row = (int64_t)blockIdx.x * rows_per_cta + threadIdx.x / LPR; row < m;
         row += (int64_t)gridDim.x * rows_per_cta) {
        const int64_t q0 = rowptr[row], q1 = rowptr[row + 1];
        E acc = zero_v(E());
        for (int64_t q = q0 + sub; q < q1; q += LPR) {
            E a = __ldg(val + q);
            if (conj_vals) a = conj_v(a);
            fmacc(acc, a, __ldg(x + __ldg(col + q)));
        }
        typename Tr<K>::W wv = widen(acc);
#pragma unroll
        for (int o = LPR / 2; o > 0; o >>= 1) {
            if constexpr (Tr<K>::cplx) {
                wv.x += __shfl_down_sync(0xffffffffu, wv.x, o, LPR);
                wv.y += __shfl_down_sync(0xffffffffu, wv.y, o, LPR);
            } else {
                wv += __shfl_down_sync(0xffffffffu, wv, o, LPR);
            }
        }
        if (sub == 0) { E r; narrow(wv, r); y[row] = r; }
    }
}

template <int K>
static void csr_t(cudaStream_t s, int64_t m, const int64_t* rowptr, const int32_t* col, const void* val,
                  const void* x, void* y, bool conj_vals, const int* flags, int sms, int lpr) {
    using E = typename Tr<K>::E;
    auto grid = [&](int L) { int64_t nb = (m * L + 255) / 256; if (nb < 1) nb = 1; if (nb > (int64_t)sms * 16) nb = (int64_t)sms * 16; return (int)nb; };
    if (lpr >= 32)      k_csr<K, 32><<<grid(32), 256, 0, s>>>(m, rowptr, col, (const E*)val, (const E*)x, (E*)y, conj_vals, flags);
    else if (lpr >= 16) k_csr<K, 16><<<grid(16), 256, 0, s>>>(m, rowptr, col, (const E*)val, (const E*)x, (E*)y, conj_vals, flags);
    else if (lpr >= 8)  k_csr<K, 8><<<grid(8), 256, 0, s>>>(m, rowptr, col, (const E*)val, (const E*)x, (E*)y, conj_vals, flags);
    else                k_csr<K, 4><<<grid(4), 256, 0, s>>>(m, rowptr, col, (const E*)val, (const E*)x, (E*)y, conj_vals, flags);
}
// lanes per row are chosen by the caller from the average row length and passed in `sms >> 16`
void launch_csr(int kind, cudaStream_t s, int64_t m, const int64_t* rowptr, const int32_t* col, const void* val,
                const void* x, void* y, bool conj_vals, const int* flags, int sms_lpr) {
    const int sms = sms_lpr & 0xffff, lpr = sms_lpr >> 16;
    switch (kind) {
        case KS: csr_t<KS>(s, m, rowptr, col, val, x, y, conj_vals, flags, sms, lpr); break;
        case KD: csr_t<KD>(s, m, rowptr, col, val, x, y, conj_vals, flags, sms, lpr); break;
        case KC: csr_t<KC>(s, m, rowptr, col, val, x, y, conj_vals, flags, sms, lpr); break;
        default: csr_t<KZ>(s, m, rowptr, col, val, x, y, conj_vals, flags, sms, lpr); break;
    }
}

// ------------------------------------------------------------------------------------------
// L2-blocked CSR SpMV (layout: lkb_csr.cu).  One launch per column block; thread-per-row (a row has ~nnz/nb entries
// in a block: 1-3 for C5), consecutive threads read consecutive table entries and nearly consecutive values.
// The x gathers of a block all fall into one L2-resident slice; they carry an L2 evict_last policy while the
// streamed operands (table, col, val, y) are marked evict_first so that they do not push the slice out.
LKB_DI uint64_t l2_policy_evict_last() { uint64_t p; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p; }
LKB_DI uint64_t l2_policy_evict_first() { uint64_t p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p; }
LKB_DI uint64_t l2_policy_evict_normal() { uint64_t p; asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p)); return p; }
template <typename T> LKB_DI T ld_hint(const T* p, uint64_t pol);
template <> LKB_DI float ld_hint<float>(const float* p, uint64_t pol) {
    float v; asm volatile("ld.global.nc.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol)); return v; }
template <> LKB_DI double ld_hint<double>(const double* p, uint64_t pol) {
    double v; asm volatile("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol)); return v; }
template <> LKB_DI float2 ld_hint<float2>(const float2* p, uint64_t pol) {
    float2 v; asm volatile("ld.global.nc.L2::cache_hint.v2.f32 {%0, %1}, [%2], %3;" : "=f"(v.x), "=f"(v.y) : "l"(p), "l"(pol)); return v; }
template <> LKB_DI double2 ld_hint<double2>(const double2* p, uint64_t pol) {
    double2 v; asm volatile("ld.global.nc.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(pol)); return v; }
template <> LKB_DI int32_t ld_hint<int32_t>(const int32_t* p, uint64_t pol) {
    int32_t v; asm volatile("ld.global.nc.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol)); return v; }
template <> LKB_DI uint32_t ld_hint<uint32_t>(const uint32_t* p, uint64_t pol) {
    uint32_t v; asm volatile("ld.global.nc.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol)); return v; }

template <int K>
__global__ void __launch_bounds__(256)
k_csr_blocked(int64_t rows, const uint32_t* __restrict__ tab, const int32_t* __restrict__ col,
              const typename Tr<K>::E* __restrict__ val, const typename Tr<K>::E* __restrict__ x,
              typename Tr<K>::E* __restrict__ y, bool conj_vals, bool first, const int* __restrict__ flags, bool stream_first)
{
    using E = typename Tr<K>::E;
    if (flags && flags[F_STOP]) return;
    const uint64_t keep = l2_policy_evict_last(), stream = stream_first ? l2_policy_evict_first() : l2_policy_evict_normal();
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t q0 = ld_hint<uint32_t>(tab + r, stream), q1 = ld_hint<uint32_t>(tab + r + 1, stream);
        if (q0 == q1 && !first) continue;                       // nothing in this block for row r
        E acc = first ? zero_v(E()) : y[r];
        for (uint32_t q = q0; q < q1; ++q) {
            E a = ld_hint<E>(val + q, stream);
            if (conj_vals) a = conj_v(a);
            fmacc(acc, a, ld_hint<E>(x + ld_hint<int32_t>(col + q, stream), keep));
        }
        y[r] = acc;
    }
}
// "CSR-stream" variant (default): a CTA owns CS_ROWS consecutive rows of the block, whose entries are ONE contiguous
// run [tab[r0], tab[r0 + CS_ROWS]) of the blocked arrays.  Phase 1: all threads stride over that run -- perfectly
// coalesced col / val streams, CS_U independent gathers of x in flight per thread -- and park the products in shared
// memory; phase 2: every thread sums the products of its rows in position order (the same fixed order as the
// thread-per-row kernel) and updates y.  The thread-per-row kernel chains tab -> col -> x -> fma with 2-3 entries per
// row and ran latency-bound (~3.3 TB/s of DRAM traffic on C5); here the three dependent loads of a chain are issued
// for CS_U entries at a time.  Streamed operands: L1::no_allocate + L2 evict_first; x gathers: L2 evict_last.
enum { CS_ROWS = 512, CS_CH = 2048 };
template <typename T> LKB_DI T ld_stream(const T* p, uint64_t pol);
template <> LKB_DI float ld_stream<float>(const float* p, uint64_t pol) {
    float v; asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol)); return v; }
template <> LKB_DI double ld_stream<double>(const double* p, uint64_t pol) {
    double v; asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol)); return v; }
template <> LKB_DI float2 ld_stream<float2>(const float2* p, uint64_t pol) {
    float2 v; asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.f32 {%0, %1}, [%2], %3;" : "=f"(v.x), "=f"(v.y) : "l"(p), "l"(pol)); return v; }
template <> LKB_DI double2 ld_stream<double2>(const double2* p, uint64_t pol) {
    double2 v; asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(pol)); return v; }
template <> LKB_DI int32_t ld_stream<int32_t>(const int32_t* p, uint64_t pol) {
    int32_t v; asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol)); return v; }
template <> LKB_DI uint32_t ld_stream<uint32_t>(const uint32_t* p, uint64_t pol) {
    uint32_t v; asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol)); return v; }

template <int K, int CS_U, int MINB>
__global__ void __launch_bounds__(256, MINB)
k_csr_blocked_stream(int64_t rows, const uint32_t* __restrict__ tab, const int32_t* __restrict__ col,
                     const typename Tr<K>::E* __restrict__ val, const typename Tr<K>::E* __restrict__ x,
                     typename Tr<K>::E* __restrict__ y, bool conj_vals, bool first, const int* __restrict__ flags)
{
    using E = typename Tr<K>::E;
    constexpr int RPT = CS_ROWS / 256;                      // rows per thread
    if (flags && flags[F_STOP]) return;
    extern __shared__ __align__(16) unsigned char cs_smem[];
    E* prod = reinterpret_cast<E*>(cs_smem);                // [CS_CH]
    const uint64_t keep = l2_policy_evict_last(), stream = l2_policy_evict_first();
    const int tid = threadIdx.x;
    for (int64_t r0 = (int64_t)blockIdx.x * CS_ROWS; r0 < rows; r0 += (int64_t)gridDim.x * CS_ROWS) {
        const int64_t rend = min(rows, r0 + CS_ROWS);
        uint32_t qa[RPT], qb[RPT];
        E acc[RPT];
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
            const int64_t r = r0 + tid + 256 * i;
            qa[i] = qb[i] = 0u; acc[i] = zero_v(E());
            if (r < rend) {
                qa[i] = ld_stream<uint32_t>(tab + r, stream); qb[i] = ld_stream<uint32_t>(tab + r + 1, stream);
                if (!first && qa[i] != qb[i]) acc[i] = y[r];
            }
        }
        const uint32_t qs = ld_stream<uint32_t>(tab + r0, stream), qe = ld_stream<uint32_t>(tab + rend, stream);
        for (uint32_t c0 = qs; c0 < qe; c0 += CS_CH) {
            const uint32_t c1 = min(c0 + (uint32_t)CS_CH, qe);
            // ---- phase 1: products of the run [c0, c1) into shared memory ----
            for (uint32_t e0 = c0 + tid; e0 < c1; e0 += 256 * CS_U) {
                int32_t cj[CS_U]; E a[CS_U], xv[CS_U];
#pragma unroll
                for (int u = 0; u < CS_U; ++u) { const uint32_t e = e0 + 256 * u; if (e < c1) cj[u] = ld_stream<int32_t>(col + e, stream); }
#pragma unroll
                for (int u = 0; u < CS_U; ++u) { const uint32_t e = e0 + 256 * u; if (e < c1) a[u] = ld_stream<E>(val + e, stream); }
#pragma unroll
                for (int u = 0; u < CS_U; ++u) { const uint32_t e = e0 + 256 * u; if (e < c1) xv[u] = ld_hint<E>(x + cj[u], keep); }
#pragma unroll
                for (int u = 0; u < CS_U; ++u) {
                    const uint32_t e = e0 + 256 * u;
                    if (e < c1) prod[e - c0] = mul_v(conj_vals ? conj_v(a[u]) : a[u], xv[u]);
                }
            }
            __syncthreads();
            // ---- phase 2: row sums in position order ----
#pragma unroll
            for (int i = 0; i < RPT; ++i) {
                const uint32_t lo = max(qa[i], c0), hi = min(qb[i], c1);
                for (uint32_t q = lo; q < hi; ++q) acc[i] = add_v(acc[i], prod[q - c0]);
            }
            __syncthreads();
        }
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
            const int64_t r = r0 + tid + 256 * i;
            if (r < rend && (first || qa[i] != qb[i])) y[r] = acc[i];
        }
    }
}
void launch_csr_blocked(int kind, cudaStream_t s, const CsrBlocked& b, const void* x, void* y, bool conj_vals, const int* flags, int sms) {
    // Variants measured on the full-size C5 matrix (B200, profiles/spmv_ab.py, gpurun_out/r02_spmv_ab*.jsonl; ncu in
    // profiles/r02_ncu_summary.md), matvec ms per sweep at 48 MB slices:
    //   0  thread-per-row, streams evict_first    17.2   DRAM 6.9 GB / block at 5.3 TB/s: bandwidth-bound on ACTUAL traffic,
    //                                                    of which ~3 GB are x-gather sectors that missed L2
    //   2  thread-per-row, streams evict_normal   16.4   (default)
    //   1  CSR-stream                             19.5   DRAM 4.9 GB / block (x misses 0.9 GB) but latency-bound at 64 regs,
    //                                                    4 CTAs / SM; forcing 5-6 CTAs / SM spills and doubles the time
    //   3  warp-pipelined CSR-stream (cp.async-staged col / val per warp, no block barrier; removed again)   20.7
    // All stream forms land at 19-21 ms whatever hides the col / val latency: the L2 gather rate (~100 G sectors/s), not the
    // streams, is what they wait for.  A persisting-L2 set-aside (64 / 79 MB) changed none of them by more than 2 %.
    static const int env_variant = getenv("LKB_CSR_BLOCKED_VARIANT") ? atoi(getenv("LKB_CSR_BLOCKED_VARIANT")) : -1;
    const int variant = env_variant >= 0 ? env_variant : b.variant;
    if (variant == 1) {
        int64_t nbs = (b.rows + CS_ROWS - 1) / CS_ROWS;
        if (nbs < 1) nbs = 1;
        if (nbs > (int64_t)sms * 6) nbs = (int64_t)sms * 6;
        const size_t sh = (size_t)CS_CH * kind_size(kind);
        for (int blk = 0; blk < b.nb; ++blk) {
            const uint32_t* tab = b.tab + (size_t)blk * (b.rows + 1);
#define LKB_CSS(K_, E_, U_, M_) k_csr_blocked_stream<K_, U_, M_><<<(int)nbs, 256, sh, s>>>(b.rows, tab, b.col, (const E_*)b.val, (const E_*)x, (E_*)y, conj_vals, blk == 0, flags)
#define LKB_CSS_K(U_, M_) switch (kind) { case KS: LKB_CSS(KS, float, U_, M_); break; case KD: LKB_CSS(KD, double, U_, M_); break; \
                                          case KC: LKB_CSS(KC, float2, U_, M_); break; default: LKB_CSS(KZ, double2, U_, M_); break; }
            LKB_CSS_K(4, 4)      // 4 gathers in flight per thread, 4 CTAs / SM (64 registers); 2 x 6 and 3 x 5 spill and run 2x slower
#undef LKB_CSS_K
#undef LKB_CSS
        }
        return;
    }
    int64_t nb = (b.rows + 255) / 256;
    if (nb < 1) nb = 1;
    if (nb > (int64_t)sms * 16) nb = (int64_t)sms * 16;
    for (int blk = 0; blk < b.nb; ++blk) {
        const uint32_t* tab = b.tab + (size_t)blk * (b.rows + 1);
        switch (kind) {
            case KS: k_csr_blocked<KS><<<(int)nb, 256, 0, s>>>(b.rows, tab, b.col, (const float*)b.val, (const float*)x, (float*)y, conj_vals, blk == 0, flags, variant == 0); break;
            case KD: k_csr_blocked<KD><<<(int)nb, 256, 0, s>>>(b.rows, tab, b.col, (const double*)b.val, (const double*)x, (double*)y, conj_vals, blk == 0, flags, variant == 0); break;
            case KC: k_csr_blocked<KC><<<(int)nb, 256, 0, s>>>(b.rows, tab, b.col, (const float2*)b.val, (const float2*)x, (float2*)y, conj_vals, blk == 0, flags, variant == 0); break;
            default: k_csr_blocked<KZ><<<(int)nb, 256, 0, s>>>(b.rows, tab, b.col, (const double2*)b.val, (const double2*)x, (double2*)y, conj_vals, blk == 0, flags, variant == 0); break;
        }
    }
}

// ------------------------------------------------------------------------------------------
// dense_linop (AbstractLinops.fypp:608-671, gemv): only the n = 128 plumbing config uses it.
template <int K>
__global__ void k_dense(int64_t m, int64_t n, const typename Tr<K>::E* __restrict__ a,
                        const typename Tr<K>::E* __restrict__ x, typename Tr<K>::E* __restrict__ y, bool trans,
                        const int* __restrict__ flags)
{
    using E = typename Tr<K>::E;
    if (flags && flags[F_STOP]) return;
    if (!trans) {
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x) {
            E s = zero_v(E());
            for (int64_t j = 0; j < n; ++j) fmacc(s, a[i + m * j], x[j]);
            y[i] = s;
        }
    } else {
        // one warp per output column
        const int lane = threadIdx.x & 31;
        for (int64_t j = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; j < n; j += ((int64_t)gridDim.x * blockDim.x) >> 5) {
            E s = zero_v(E());
            for (int64_t i = lane; i < m; i += 32) fma_conj(s, a[i + m * j], x[i]);
            typename Tr<K>::W wv = warp_sum(widen(s));
            if (lane == 0) { E r; narrow(wv, r); y[j] = r; }
        }
    }
}
void launch_dense(int kind, cudaStream_t s, int64_t m, int64_t n, const void* a, const void* x, void* y, bool trans,
                  const int* flags) {
    const int64_t work = trans ? n * 32 : m;
    int nb = (int)((work + 127) / 128); if (nb < 1) nb = 1; if (nb > 4096) nb = 4096;
    switch (kind) {
        case KS: k_dense<KS><<<nb, 128, 0, s>>>(m, n, (const float*)a, (const float*)x, (float*)y, trans, flags); break;
        case KD: k_dense<KD><<<nb, 128, 0, s>>>(m, n, (const double*)a, (const double*)x, (double*)y, trans, flags); break;
        case KC: k_dense<KC><<<nb, 128, 0, s>>>(m, n, (const float2*)a, (const float2*)x, (float2*)y, trans, flags); break;
        default: k_dense<KZ><<<nb, 128, 0, s>>>(m, n, (const double2*)a, (const double2*)x, (double2*)y, trans, flags); break;
    }
}

}  // namespace lkb
