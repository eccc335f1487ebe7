// lkb_step.cuh -- the O(1) decisions of one Krylov step on the new sub-diagonal entry beta = ||w||, shared by the
// 1-CTA update kernel (k_update) and by the fused final multi-axpy (k_multiaxpy_fin).
// Reference: src/Krylov/qr.fypp:137-164 (qr_no_pivoting, p = 1), arnoldi.fypp:59-71, lanczos.fypp:29-40,
// golub_kahan.fypp:37-59.
#pragma once
#include "lkb_kernels.h"

namespace lkb {

// decisions of qr_no_pivoting / arnoldi / lanczos / bidiagonalization on the new sub-diagonal entry; one thread
template <int K>
LKB_DI void step_decide(double beta, typename Tr<K>::E* hcol, int j, double tol, double atol, double* inv_dev,
                        int* flags, int kstep, int mode)
{
    using E = typename Tr<K>::E;
    if (beta != beta) flags[F_NAN] = 1;
    Scalar hb; hb.re = beta; hb.im = 0.0;
    double inv = 1.0;
    if (mode == 0) {
        // qr_no_pivoting (p = 1) uses its own default tol = atol for the refill decision
        // (qr.fypp:125,146), arnoldi then tests |H(k+1,k)| < tol (arnoldi.fypp:59-71).
        double hkk = beta;
        if (beta < atol) { hkk = 0.0; flags[F_REFILL] = 1; }
        else inv = 1.0 / beta;
        hb.re = hkk;
        if (hkk < tol) { flags[F_STOP] = 1; flags[F_INFO] = kstep; }
    } else if (mode == 1) {          // lanczos.fypp:29-40: beta < tol => exit, no scaling
        if (beta < tol) { flags[F_STOP] = 1; flags[F_INFO] = kstep; flags[F_REFILL] = 1; }
        else inv = 1.0 / beta;
    } else if (mode == 2) {          // golub_kahan.fypp:37-43: scale iff beta > tol
        if (beta > tol) inv = 1.0 / beta;
        else { flags[F_STOP] = 1; flags[F_INFO] = kstep; flags[F_REFILL] = 1; }
    } else {                         // plain norm: no decision
        inv = beta > 0.0 ? 1.0 / beta : 1.0;
    }
    if (hcol) { E h; from_scalar(hb, h); hcol[j] = h; }
    *inv_dev = inv;
}
// the scaling rule of each mode, evaluated by every CTA of the fused kernel on the predicted beta
LKB_DI double fin_inv(int mode, double beta, double tol, double atol) {
    if (mode == 0) return (beta < atol) ? 1.0 : 1.0 / beta;
    if (mode == 1) return (beta < tol) ? 1.0 : 1.0 / beta;
    return (beta > tol) ? 1.0 / beta : 1.0;          // bidiag, gmres
}


}  // namespace lkb
