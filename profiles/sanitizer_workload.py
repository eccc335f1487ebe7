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
    # round 2: blocked CSR layout (forced on a small matrix), bidiagonalisation through it, basis GEMM (DMMA for rdp)
    if kind != "s":
        rng = np.random.default_rng(5)
        m, nn, pr = 3001, 2003, 6
        rp = np.arange(0, (m + 1) * pr, pr, dtype=np.int64)
        ci = np.sort(rng.integers(0, nn, size=(m, pr), dtype=np.int32), axis=1).ravel()
        va = rng.standard_normal(m * pr).astype(lk.DTYPES[kind])
        ctx.set_option("csr_slice_kb", 4); ctx.set_option("csr_block_min_kb", 0)
        Ac = lk.LinOp.csr(ctx, m, nn, rp, ci, va)
        ctx.set_option("csr_slice_kb", 48 * 1024); ctx.set_option("csr_block_min_kb", 96 * 1024)
        U = lk.Basis(ctx, kind, m, 9); V = lk.Basis(ctx, kind, nn, 9)
        u0 = U.col(0).fill_random("normal", 47); u0.scal(1.0 / u0.norm())
        B = np.zeros((9, 8), dtype=lk.DTYPES[kind], order="F")
        assert lk.bidiagonalization(Ac, U, V, B) == 0
        lk.set_lapack_from_scipy()
        Hs = np.asfortranarray(np.triu(rng.standard_normal((kdim + 1, kdim)), -1).astype(lk.DTYPES[kind]))
        lk.krylov_schur(X, Hs, kdim)
    print(kind, info, g)
print("sanity done")
