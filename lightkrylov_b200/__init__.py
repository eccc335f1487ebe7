"""lightkrylov_b200 -- B200-native Krylov-factorisation hot path behind LightKrylov's API.

The package is a thin host mirror (ctypes) of the C ABI in include/lkb.h; all arithmetic runs
in hand-written sm_100a CUDA kernels inside csrc/liblkb.so.  There is no CPU fallback.
"""
from ._lib import LkbError, SO_PATH, load  # noqa: F401
from .api import (ATOL, DTYPES, KINDS, RTOL, Basis, Context, LinOp, Vector, arnoldi, bidiagonalization, cg,  # noqa: F401
                  double_gram_schmidt_step, eighs, eigs, fgmres, gmres, initialize_krylov_subspace,
                  initialize_random_orthonormal_basis, kexpm, kexpm_mat, kind_of, krylov_exptA, krylov_schur, lanczos,
                  orthogonalize_against_basis, partition, qr, qr_pivoting, save_eigenspectrum, set_lapack_from_scipy, svds,
                  write_results)
