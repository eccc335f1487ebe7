// lkb_eig.cu -- eigs / eighs / svds / krylov_schur host shells (placeholder: filled in below)
#include "../../include/lkb.h"
#include "lkb_internal.h"
using namespace lkb;
extern "C" {
int lkb_set_lapack(const char*, const char*, const char*) { set_error("lkb_set_lapack: not implemented yet"); return LKB_ERR_LAPACK; }
int lkb_krylov_schur(lkb_basis_t, void*, int, int, int32_t*) { set_error("krylov_schur: not implemented yet"); return LKB_ERR_LAPACK; }
int lkb_eigs(lkb_op_t, lkb_basis_t, int, double*, double*, int32_t*, lkb_vec_t, int32_t, double, int32_t) { set_error("eigs: not implemented yet"); return LKB_ERR_LAPACK; }
int lkb_eighs(lkb_op_t, lkb_basis_t, int, double*, double*, int32_t*, lkb_vec_t, int32_t, double) { set_error("eighs: not implemented yet"); return LKB_ERR_LAPACK; }
int lkb_svds(lkb_op_t, lkb_basis_t, double*, lkb_basis_t, int, double*, int32_t*, lkb_vec_t, int32_t, double) { set_error("svds: not implemented yet"); return LKB_ERR_LAPACK; }
}
