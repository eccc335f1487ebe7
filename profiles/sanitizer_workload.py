import sys; sys.path.insert(0, ".")
import numpy as np, lightkrylov_b200 as lk
ctx = lk.Context(0)
for kind in ("d", "z", "s"):
    nx, ny, kdim = 96, 70, 40; n = nx * ny
    A = lk.LinOp.stencil5(ctx, kind, nx, ny, (4.0, -1.0, -1.0, -1.0, -1.0))
    X = lk.Basis(ctx, kind, n, kdim + 1)
    x0 = X.col(0).fill_random("uniform", 42); x0.scal(1.0 / x0.norm())
    H = np.zeros((kdim + 1, kdim), dtype=lk.DTYPES[kind], order="F")
    assert lk.arnoldi(A, X, H) == 0
    T = np.zeros_like(H)
    assert lk.lanczos(A, X, T) == 0
    A3 = lk.LinOp.stencil7(ctx, kind, 20, 16, 12, (6.0, -1.0, -1.0, -1.0, -1.0, -1.0, -1.0))
    b = lk.Vector(ctx, kind, 20 * 16 * 12).fill_random("uniform", 1); x = lk.Vector(ctx, kind, 20 * 16 * 12)
    info, meta = lk.cg(A3, b, x, maxiter=300)
    g, gm = lk.gmres(A3, b, x.zero(), kdim=20, maxiter=5)
    print(kind, info, g)
print("sanity done")
