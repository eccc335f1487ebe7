#!/usr/bin/env python
"""BASELINE configs[1] at FULL size (5-point Poisson 4096^2, n = 16.8M, fp64, arnoldi kdim = 128, x0 = U[0,1) seed 42 normalised --
the configuration of bench.py and of c2_full_H.npz) computed by THE REFERENCE'S OWN arnoldi / Gram-Schmidt / qr sources, executed by
oracle/f90run.py on the reference's dense_vector_rdp with the user-side stencil operator of user_stencil.f90.

    python tests/golden/make_ref_golden_c2.py [kdim]        # container only (needs /root/reference); ~30-40 min, ~20 GB of host memory

Writes ref_c2_full_H.npz (H, info, Ritz values, head of the last basis vector).  tests/test_ref_golden.py compares it with the
oracle's c2_full_H.npz -- the matrix every bench.py line is checked against -- so the headline parity record is pinned to the
reference's code at the headline size.  For this one run the three BLAS-1 natives of the interpreter (stdlib's dot / axpy / scal, which
the reference's dense_vector calls) are served by OpenBLAS instead of the interpreter's left-to-right numpy loops: 16.8M-element vectors,
33 000 dots and axpys -- the summation order inside a dot product is therefore OpenBLAS' (the comparison tolerance is 1e-12).
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)

if __name__ == "__main__":
    from scipy.linalg import blas
    from oracle import f90run, lk_oracle as lo, ref_exec
    import ref_cases as rc
    kdim = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    nx = ny = int(os.environ.get("LK_C2_N", "4096"))
    n = nx * ny
    it = ref_exec.interp()
    it.p.load(os.path.join(HERE, "user_stencil.f90"))
    f90run.Interp(it.p)

    def n_dot(interp, m, x, incx, y, incy):
        return np.float64(blas.ddot(x, y, n=int(m)))

    def n_axpy(interp, m, a, x, incx, y, incy):
        out = blas.daxpy(x, y, n=int(m), a=float(a))
        assert out is y or np.shares_memory(out, y)

    def n_scal(interp, m, a, x, incx):
        out = blas.dscal(float(a), x, n=int(m))
        assert out is x or np.shares_memory(out, x)
    it.natives.update({"dot": n_dot, "axpy": n_axpy, "scal": n_scal})

    op = it.new_inst("stencil_linop_rdp")
    op.f["nx"], op.f["ny"], op.f["nz"] = nx, ny, 1
    op.f["coef"][:5] = rc.POISSON2D
    x0 = lo.fill(n, "d", "uniform", 42)
    lo.normalize(x0)
    X = np.empty(kdim + 1, dtype=object)
    for i in range(kdim + 1):
        X[i] = it.new_inst("dense_vector_rdp")
        X[i].f["n"] = n
        X[i].f["data"] = x0.copy() if i == 0 else np.zeros(n)
    H = np.zeros((kdim + 1, kdim), order="F")
    t0 = time.time()
    _, o = it.call("arnoldi", op, X, H, 0)
    dt = time.time() - t0
    ritz = np.sort(np.linalg.eigvals(H[:kdim, :kdim]).real)
    tmp = os.path.join(HERE, "ref_c2_full_H.tmp.npz")
    np.savez_compressed(tmp, H=H, info=int(o[3]), ritz=ritz, seconds=dt, nx=nx, kdim=kdim, x0_head=X[0].f["data"][:8].copy(),
                        xlast_head=X[kdim].f["data"][:8].copy(), matvecs=int(op.f["matvec_counter"]))
    os.replace(tmp, os.path.join(HERE, "ref_c2_full_H.npz" if (nx, kdim) == (4096, 128) else f"ref_c2_{nx}_{kdim}_H.npz"))
    print("reference C2 written: info", int(o[3]), "seconds", dt)
