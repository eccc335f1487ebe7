"""Small workloads for ncu captures of the non-headline kernels (3-D stencil, CG iteration, basis GEMM).

    ncu --set full -k regex:<kernel> ... python profiles/ncu_targets.py stencil3d|cg|gemm
"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lightkrylov_b200 as lk

what = sys.argv[1]
ctx = lk.Context(0)
LAP7 = (6., -1., -1., -1., -1., -1., -1.)
if what == "stencil3d":
    m = 384
    A = lk.LinOp.stencil7(ctx, "d", m, m, m, LAP7)
    x = lk.Vector(ctx, "d", m ** 3).fill_random("uniform", 1); y = lk.Vector(ctx, "d", m ** 3)
    for _ in range(8):
        A.matvec(x, y); A.matvec(y, x)
elif what == "cg":
    m = 384
    A = lk.LinOp.stencil7(ctx, "d", m, m, m, LAP7)
    b = lk.Vector(ctx, "d", m ** 3).fill_random("uniform", 45); x = lk.Vector(ctx, "d", m ** 3)
    ctx.set_option("graphs", 0)
    lk.cg(A, b, x, maxiter=12)
elif what == "gemm":
    n, k, p = 16 * 1024 * 1024, 128, 64
    X = lk.Basis(ctx, "d", n, k + 1)
    for i in range(4):
        X.col(i).fill_random("normal", i + 1)
    lk.set_lapack_from_scipy()
    H = np.triu(np.random.default_rng(0).standard_normal((k + 1, k)), -1)
    H = np.asfortranarray(H)
    lk.krylov_schur(X, H, k)
ctx.sync()
