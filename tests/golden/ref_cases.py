"""Cases shared by tests/golden/make_ref_golden.py (runs THE REFERENCE'S SOURCES through oracle/f90run.py, writes the fixtures)
and tests/test_ref_golden.py (re-creates the same inputs, runs the C oracle -- and, on a GPU, the product -- and compares with
the committed fixtures).  Inputs come from `pseudo` (integer hashing, bit-reproducible); only OUTPUTS are stored.

Every case is a function  case(kind, be) -> dict of output arrays,  written once against a tiny backend interface `be`
(RefBackend below = the reference under the interpreter; the tests supply an oracle backend and a product backend), so the
three implementations are driven with literally the same calls and inputs.
"""
import numpy as np

N = 128                      # the reference's test_size (src/Utilities/TestUtils.f90:16)
DTYPE = {"s": np.float32, "d": np.float64, "c": np.complex64, "z": np.complex128}
KINDS = "sdcz"


def pseudo(shape, seed, kind="d"):
    """same generator as oracle/ref_exec.pseudo (duplicated so that the tests do not need oracle/ref_exec's dependencies)"""
    n = int(np.prod(shape))

    def one(sd):
        x = (np.arange(n, dtype=np.uint64) + np.uint64(sd) * np.uint64(0x9E3779B1)) & np.uint64(0xFFFFFFFF)
        x = (x * np.uint64(2654435761)) & np.uint64(0xFFFFFFFF)
        x ^= x >> np.uint64(16)
        x = (x * np.uint64(2246822519)) & np.uint64(0xFFFFFFFF)
        x ^= x >> np.uint64(13)
        x = (x * np.uint64(3266489917)) & np.uint64(0xFFFFFFFF)
        x ^= x >> np.uint64(16)
        return ((x >> np.uint64(8)).astype(np.float64) / float(1 << 24) - 0.5)
    v = one(seed)
    if kind in "cz":
        v = v + 1j * one(seed + 7919)
    return np.asfortranarray(v.astype(DTYPE[kind]).reshape(shape, order="F"))


def unit(v):
    """normalised in fp64, then rounded to the kind (so every implementation starts from identical bits)"""
    w = v.astype(np.complex128 if np.iscomplexobj(v) else np.float64)
    return (w / np.linalg.norm(w)).astype(v.dtype)


def orthonormal_block(kind, ncols, seed):
    M = pseudo((N, ncols), seed, kind)
    Q, _ = np.linalg.qr(M.astype(np.complex128 if kind in "cz" else np.float64))
    return np.asfortranarray(Q.astype(DTYPE[kind]))


def general_matrix(kind, seed=1):
    return pseudo((N, N), seed, kind)                       # entries in [-0.5, 0.5): spectral radius ~ 3.3


def sym_matrix(kind, seed=3, shift=0.01):
    M = pseudo((N, N), seed, kind).astype(np.complex128 if kind in "cz" else np.float64)
    S = M @ M.conj().T / N + shift * np.eye(N)
    return np.asfortranarray(((S + S.conj().T) / 2).astype(DTYPE[kind]))


def block_triangular(kind, m, seed=5):
    """A = [[B, C], [0, D]] with B m x m: a start vector supported on the first m entries spans an invariant subspace of
    dimension m, so the factorisations break down (info = m) at step m."""
    A = pseudo((N, N), seed, kind)
    A[m:, :m] = 0
    return A


# ------------------------------------------------------------------------------------------------------------------ cases
def arnoldi_full(kind, be):
    kdim = 24
    A = be.linop(kind, general_matrix(kind))
    X = be.basis(kind, kdim + 1, unit(pseudo((N,), 2, kind)))
    H = np.zeros((kdim + 1, kdim), dtype=DTYPE[kind], order="F")
    info = be.arnoldi(A, X, H)
    return {"info": info, "H": H, "X": be.data(X), "matvecs": be.counter(A)}


def arnoldi_transpose(kind, be):
    kdim = 10
    A = be.linop(kind, general_matrix(kind, 11))
    X = be.basis(kind, kdim + 1, unit(pseudo((N,), 12, kind)))
    H = np.zeros((kdim + 1, kdim), dtype=DTYPE[kind], order="F")
    info = be.arnoldi(A, X, H, transpose=True)
    return {"info": info, "H": H, "X": be.data(X)}


def arnoldi_block(kind, be):
    p, kdim = 2, 8
    A = be.linop(kind, general_matrix(kind, 21))
    X = be.basis(kind, p * (kdim + 1), orthonormal_block(kind, p, 22))
    H = np.zeros((p * (kdim + 1), p * kdim), dtype=DTYPE[kind], order="F")
    info = be.arnoldi(A, X, H, blksize=p)
    return {"info": info, "H": H, "X": be.data(X)}


def arnoldi_resume(kind, be):
    kdim = 12
    A = be.linop(kind, general_matrix(kind, 31))
    X = be.basis(kind, kdim + 1, unit(pseudo((N,), 32, kind)))
    H = np.zeros((kdim + 1, kdim), dtype=DTYPE[kind], order="F")
    info1 = be.arnoldi(A, X, H, kend=6)
    H6 = H.copy()
    info2 = be.arnoldi(A, X, H, kstart=7, kend=12)
    return {"info1": info1, "info2": info2, "H_after_6": H6, "H": H, "X": be.data(X)}


def arnoldi_breakdown(kind, be):
    kdim, m = 10, 4
    A = be.linop(kind, block_triangular(kind, m))
    x0 = pseudo((N,), 42, kind)
    x0[m:] = 0
    X = be.basis(kind, kdim + 1, unit(x0))
    H = np.zeros((kdim + 1, kdim), dtype=DTYPE[kind], order="F")
    tol = 1e-4 if kind in "sc" else 1e-10
    info = be.arnoldi(A, X, H, tol=tol)
    Xd = be.data(X)
    return {"info": info, "H_lead": H[:m, :m].copy(), "X_lead": Xd[:, :m].copy()}


def lanczos_full(kind, be):
    kdim = 24
    A = be.linop(kind, sym_matrix(kind), sym=True)
    X = be.basis(kind, kdim + 1, unit(pseudo((N,), 52, kind)))
    T = np.zeros((kdim + 1, kdim), dtype=DTYPE[kind], order="F")
    info = be.lanczos(A, X, T)
    return {"info": info, "T": T, "X": be.data(X)}


def lanczos_resume(kind, be):
    """kstart / kend: five steps, then the rest -- same T and basis as one call (lanczos.fypp:7-64)"""
    kdim = 12
    A = be.linop(kind, sym_matrix(kind, 53), sym=True)
    X = be.basis(kind, kdim + 1, unit(pseudo((N,), 54, kind)))
    T = np.zeros((kdim + 1, kdim), dtype=DTYPE[kind], order="F")
    info1 = be.lanczos(A, X, T, kend=5)
    T5 = T.copy()
    info2 = be.lanczos(A, X, T, kstart=6, kend=12)
    return {"info1": info1, "info2": info2, "T_after_5": T5, "T": T, "X": be.data(X)}


def bidiag_resume(kind, be):
    kdim = 10
    A = be.linop(kind, general_matrix(kind, 63))
    U = be.basis(kind, kdim + 1, unit(pseudo((N,), 64, kind)))
    V = be.basis(kind, kdim + 1)
    B = np.zeros((kdim + 1, kdim), dtype=DTYPE[kind], order="F")
    info1 = be.bidiag(A, U, V, B, kend=4)
    B4 = B.copy()
    info2 = be.bidiag(A, U, V, B, kstart=5, kend=10)
    return {"info1": info1, "info2": info2, "B_after_4": B4, "B": B, "U": be.data(U), "V": be.data(V)}


def bidiag_full(kind, be):
    kdim = 16
    A = be.linop(kind, general_matrix(kind, 61))
    U = be.basis(kind, kdim + 1, unit(pseudo((N,), 62, kind)))
    V = be.basis(kind, kdim + 1)
    B = np.zeros((kdim + 1, kdim), dtype=DTYPE[kind], order="F")
    info = be.bidiag(A, U, V, B)
    return {"info": info, "B": B, "U": be.data(U), "V": be.data(V)}


def qr_full(kind, be):
    p = 8
    Q = be.basis(kind, p, pseudo((N, p), 71, kind))
    info, R = be.qr(Q)
    return {"info": info, "R": R, "Q": be.data(Q)}


def qr_deficient(kind, be):
    """column 4 is a combination of columns 1..3: info = 4, R(4,4) = 0, the column is refilled with random numbers (whose
    generator differs between implementations), so only what precedes the refill is comparable"""
    p = 6
    M = pseudo((N, p), 81, kind)
    M[:, 3] = (0.5 * M[:, 0] - 0.25 * M[:, 1] + 2.0 * M[:, 2]).astype(DTYPE[kind])
    Q = be.basis(kind, p, M)
    tol = 1e-4 if kind in "sc" else 1e-10
    info, R = be.qr(Q, tol=tol)
    Qd = be.data(Q)
    G = Qd.conj().T @ Qd
    return {"info": info, "R_lead": R[:4, :4].copy(), "Q_lead": Qd[:, :3].copy(),
            "abs_orth_err": np.array(np.abs(G - np.eye(p)).max(), dtype=np.float64)}      # key prefix abs_: bounded, not compared


def qr_pivoting(kind, be):
    p = 8
    M = pseudo((N, p), 91, kind) * (1.0 + np.arange(p))[None, :].astype(DTYPE[kind])       # distinct column norms: unambiguous pivots
    Q = be.basis(kind, p, M.astype(DTYPE[kind]))
    info, R, perm = be.qr_pivoting(Q)
    return {"info": info, "R": R, "perm": np.asarray(perm, dtype=np.int64), "Q": be.data(Q)}


def dgs_vector(kind, be):
    j = 6
    X = be.basis(kind, j, orthonormal_block(kind, j, 101))
    y = be.basis(kind, 1, pseudo((N,), 102, kind))
    info, beta = be.dgs_vec(y, X)
    return {"info": info, "beta": beta, "y": be.data(y)[:, 0].copy()}


def dgs_basis(kind, be):
    j, p = 6, 3
    X = be.basis(kind, j, orthonormal_block(kind, j, 111))
    Y = be.basis(kind, p, pseudo((N, p), 112, kind))
    info, beta = be.dgs_bas(Y, X)
    return {"info": info, "beta": beta, "Y": be.data(Y)}


def dgs_zero_vector(kind, be):
    """orthogonalize_against_basis flags a vector whose norm is below atol (info = 1) -- the `info` of the Gram-Schmidt passes"""
    j = 4
    X = be.basis(kind, j, orthonormal_block(kind, j, 121))
    y = be.basis(kind, 1)                                   # the zero vector
    info, beta = be.dgs_vec(y, X)
    return {"info": info, "beta": beta, "y": be.data(y)[:, 0].copy()}


def qr_pivoting_deficient(kind, be):
    """exact rank deficiency (cf. test_pivoting_qr_exact_rank_deficiency, test/TestKrylov.fypp): columns 3 and 6 of 6 are zero:
    the pivoting order, info and the leading block of R are defined by the data; the refilled columns are random"""
    p = 6
    M = pseudo((N, p), 131, kind) * (1.0 + np.arange(p))[None, :].astype(DTYPE[kind])
    M[:, 2] = 0
    M[:, 5] = 0
    Q = be.basis(kind, p, M.astype(DTYPE[kind]))
    info, R, perm = be.qr_pivoting(Q)
    Qd = be.data(Q)
    G = Qd.conj().T @ Qd
    return {"info": info, "R_lead": R[:4, :4].copy(), "perm_lead": np.asarray(perm, dtype=np.int64)[:4].copy(),
            "Q_lead": Qd[:, :4].copy(), "abs_orth_err": np.array(np.abs(G - np.eye(p)).max(), dtype=np.float64)}


def krylov_schur_restart(kind, be):
    """a full Arnoldi factorisation followed by ONE Krylov-Schur restart with eigs' median selector (BaseKrylov.fypp:782-834):
    number of retained Ritz values, restarted H (Schur form + residual row), restarted basis"""
    kdim = 16
    A = be.linop(kind, dominant_matrix(kind))
    X = be.basis(kind, kdim + 1, unit(pseudo((N,), 502, kind)))
    H = np.zeros((kdim + 1, kdim), dtype=DTYPE[kind], order="F")
    info = be.arnoldi(A, X, H)
    n = be.krylov_schur(X, H)
    Xd = be.data(X)
    # the Schur vectors of a (quasi-)triangular form are defined up to a sign / phase per block: compare what is invariant
    # ... and, for the real kinds, up to the orientation of each standardised 2 x 2 block ((q1, q2) -> (q2, -q1) is the same form):
    # sorted moduli are invariant under both
    return {"info": info, "n": n, "absH_diag": np.sort(np.abs(np.diag(H[:n, :n]))), "absres_row": np.sort(np.abs(H[n, :n])),
            "H_fro": np.array(np.linalg.norm(H), dtype=np.float64),
            "span_projector_diag": np.real(np.einsum("ij,ij->i", Xd[:, :n + 1], Xd[:, :n + 1].conj())).astype(np.float64)}


def basis_helpers(kind, be):
    """innerprod (vector and matrix), Gram, linear_combination (vector and matrix), axpby_basis, copy: AbstractVectors.fypp:571-730"""
    j, p = 5, 3
    Xh = pseudo((N, j), 511, kind)
    Yh = pseudo((N, p), 512, kind)
    Bm = pseudo((j, p), 513, kind)
    X, Y = be.basis(kind, j, Xh), be.basis(kind, p, Yh)
    out = be.helpers(kind, X, Y, Bm)
    return out


CASES = {
    "lanczos_resume": lanczos_resume, "bidiag_resume": bidiag_resume, "krylov_schur_restart": krylov_schur_restart, "basis_helpers": basis_helpers,
    "dgs_zero_vector": dgs_zero_vector, "qr_pivoting_deficient": qr_pivoting_deficient,
    "arnoldi_full": arnoldi_full, "arnoldi_transpose": arnoldi_transpose, "arnoldi_block": arnoldi_block,
    "arnoldi_resume": arnoldi_resume, "arnoldi_breakdown": arnoldi_breakdown, "lanczos_full": lanczos_full,
    "bidiag_full": bidiag_full, "qr_full": qr_full, "qr_deficient": qr_deficient, "qr_pivoting": qr_pivoting,
    "dgs_vector": dgs_vector, "dgs_basis": dgs_basis,
}



# ------------------------------------------------------------------------------------------------------------------ solvers
def _real(kind):
    return np.float32 if kind in "sc" else np.float64


def well_conditioned(kind, seed=201):
    """A = 1.5 I + M * 2 / sqrt(N) (entries of M in [-0.5, 0.5)): gmres(10) needs several restarts"""
    M = pseudo((N, N), seed, kind)
    A = (1.5 * np.eye(N) + M * (2.0 / np.sqrt(N))).astype(DTYPE[kind])
    return np.asfortranarray(A)


def gmres_solve(kind, be):
    A = be.linop(kind, well_conditioned(kind))
    b = be.basis(kind, 1, unit(pseudo((N,), 202, kind)))
    x = be.basis(kind, 1)
    info, meta = be.gmres(A, b, x, kdim=10, maxiter=20)
    return {"info": info, "x": be.data(x)[:, 0].copy(), "res": np.asarray(meta["res"], dtype=np.float64),
            "n_iter": meta["n_iter"], "n_inner": meta["n_inner"], "n_outer": meta["n_outer"]}


def cg_solve(kind, be):
    # shift 0.1: condition number ~ 4: the CG recurrence (not backward stable) stays within rounding of itself across
    # implementations for the whole history; with the 0.01 shift of the reference's own test operator the iterates of two
    # implementations drift apart by 20 % after 45 steps and the iteration counts differ by one
    A = be.linop(kind, sym_matrix(kind, 211, shift=0.1), sym=True)
    b = be.basis(kind, 1, unit(pseudo((N,), 212, kind)))
    x = be.basis(kind, 1)
    info, meta = be.cg(A, b, x, maxiter=300)
    return {"info": info, "x": be.data(x)[:, 0].copy(), "res": np.asarray(meta["res"], dtype=np.float64),
            "n_iter": meta["n_iter"]}


def eighs_solve(kind, be):
    nev = 4
    A = be.linop(kind, sym_matrix(kind, 221), sym=True)
    x0 = unit(pseudo((N,), 222, kind))
    ev, res, X, info = be.eighs(A, nev, x0, kdim=64, tolerance=1e-3 if kind in "sc" else 1e-9)
    return {"info": info, "eigvals": np.asarray(ev, dtype=np.float64), "absvecs": np.abs(X)}       # eigenvectors up to a phase


def svds_solve(kind, be):
    nsv = 3
    A = be.linop(kind, general_matrix(kind, 231))
    u0 = unit(pseudo((N,), 232, kind))
    S, res, U, V, info = be.svds(A, nsv, u0, kdim=64, tolerance=1e-3 if kind in "sc" else 1e-9)
    return {"info": info, "S": np.asarray(S, dtype=np.float64), "absU": np.abs(U), "absV": np.abs(V)}


def dominant_matrix(kind, seed=251):
    """non-normal matrix with four well separated leading eigenvalues (the rest inside a disc of radius ~ 0.6)"""
    A = pseudo((N, N), seed, kind) * DTYPE[kind](2.0 / np.sqrt(N))
    for i, d in enumerate((3.0, 2.5, -2.2, 1.9)):
        A[i, i] += DTYPE[kind](d)
    return np.asfortranarray(A.astype(DTYPE[kind]))


def eigs_solve(kind, be):
    nev = 4
    A = be.linop(kind, dominant_matrix(kind))
    x0 = unit(pseudo((N,), 252, kind))
    ev, res, X, info = be.eigs(A, nev, x0, kdim=24, tolerance=1e-3 if kind in "sc" else 1e-9)
    out = {"info": info, "eigvals": np.asarray(ev, dtype=np.complex128)}
    if kind in "sd":
        out["X"] = X                      # real kinds: (Re, Im) column pairs, normalised by xGEEV -- directly comparable
    else:
        out["absX"] = np.abs(X)           # complex kinds: eigenvectors are defined up to a phase
    return out


def kexpm_solve(kind, be):
    A = be.linop(kind, general_matrix(kind, 241))
    b = be.basis(kind, 1, unit(pseudo((N,), 242, kind)))
    c = be.basis(kind, 1)
    info = be.kexpm(c, A, b, tau=0.1, tol=1e-4 if kind in "sc" else 1e-10, kdim=40)
    return {"info": info, "c": be.data(c)[:, 0].copy()}


def eighs_write_intermediate(kind, be):
    """write_results sorts its residual argument in place (IterativeSolvers.fypp:882-924): with write_intermediate the
    residuals eighs returns are the LARGEST entries of the ascending table -- a literal side effect of the reference"""
    nev = 4
    A = be.linop(kind, sym_matrix(kind, 221), sym=True)
    x0 = unit(pseudo((N,), 222, kind))
    ev, res, X, info = be.eighs(A, nev, x0, kdim=64, tolerance=1e-3 if kind in "sc" else 1e-9, write_intermediate=True)
    return {"info": info, "eigvals": np.asarray(ev, dtype=np.float64), "residuals": np.asarray(res, dtype=np.float64)}


def svds_write_intermediate(kind, be):
    nsv = 3
    A = be.linop(kind, general_matrix(kind, 231))
    u0 = unit(pseudo((N,), 232, kind))
    S, res, U, V, info = be.svds(A, nsv, u0, kdim=64, tolerance=1e-3 if kind in "sc" else 1e-9, write_intermediate=True)
    # the residuals svds returns are then the SMALLEST entries of the ascending table (~ tolerance): converged quantities
    # with no relative accuracy, so only their bound is checked
    return {"info": info, "S": np.asarray(S, dtype=np.float64),
            "abs_residual_max": np.array(np.abs(res).max() * (1e-9 if kind in "dz" else 1e-3), dtype=np.float64)}


def fgmres_solve(kind, be):
    A = be.linop(kind, well_conditioned(kind, 261))
    b = be.basis(kind, 1, unit(pseudo((N,), 262, kind)))
    x = be.basis(kind, 1)
    info, meta = be.gmres(A, b, x, kdim=8, maxiter=30, flexible=True)
    return {"info": info, "x": be.data(x)[:, 0].copy(), "res": np.asarray(meta["res"], dtype=np.float64),
            "n_iter": meta["n_iter"], "n_inner": meta["n_inner"], "n_outer": meta["n_outer"]}


def kexpm_breakdown(kind, be):
    """the Arnoldi factorisation inside kexpm breaks down (invariant subspace of dimension 4): literal `info` and result"""
    m = 4
    A = be.linop(kind, block_triangular(kind, m, seed=281))
    b0 = pseudo((N,), 282, kind)
    b0[m:] = 0
    b = be.basis(kind, 1, unit(b0))
    c = be.basis(kind, 1)
    info = be.kexpm(c, A, b, tau=0.2, tol=1e-4 if kind in "sc" else 1e-10, kdim=20)
    return {"info": info, "c": be.data(c)[:, 0].copy()}


def kexpm_block(kind, be):
    p = 3
    A = be.linop(kind, general_matrix(kind, 271))
    B = be.basis(kind, p, pseudo((N, p), 272, kind))
    Cb = be.basis(kind, p)
    info = be.kexpm_mat(Cb, A, B, tau=0.1, tol=1e-4 if kind in "sc" else 1e-10, kdim=20)
    return {"info": info, "C": be.data(Cb)}


# ---- the north-star operator family (BASELINE configs C2 / C3 / C4 at reduced size): constant-coefficient stencils written as
# USER code against the reference's abstract types (tests/golden/user_stencil.f90), real(dp) only
POISSON2D = (4.0, -1.0, -1.0, -1.0, -1.0)
CONVDIFF2D = (6.0, -1.3, -0.7, -1.2, -0.8)
POISSON3D = (6.0, -1.0, -1.0, -1.0, -1.0, -1.0, -1.0)
CONVDIFF3D = (6.0, -1.4, -0.6, -1.3, -0.7, -1.2, -0.8)


def stencil2d_arnoldi(kind, be):
    """C2: arnoldi on the 5-point Poisson operator (48 x 40 here; 4096 x 4096 in the bench)"""
    dims, kdim = (48, 40), 32
    n = dims[0] * dims[1]
    A = be.stencil(kind, dims, POISSON2D)
    X = be.basis_n(kind, n, kdim + 1, unit(pseudo((n,), 402, kind)))
    H = np.zeros((kdim + 1, kdim), dtype=DTYPE[kind], order="F")
    info = be.arnoldi(A, X, H)
    return {"info": info, "H": H, "X": be.data(X), "matvecs": be.counter(A)}


def stencil2d_arnoldi_large(kind, be):
    """C2 at a size where the product's tiling is fully engaged (n = 786432 rows, 64 steps: many CTAs per kernel, j > 16 so that
    the staged TMA ring and all chunk warps of the fused Gram-Schmidt kernel are live): every Hessenberg entry against the
    reference's own arnoldi.  Only H, the head of the last basis vector and an orthonormality figure are stored."""
    dims, kdim = (1024, 768), 64
    n = dims[0] * dims[1]
    A = be.stencil(kind, dims, POISSON2D)
    X = be.basis_n(kind, n, kdim + 1, unit(pseudo((n,), 452, kind)))
    H = np.zeros((kdim + 1, kdim), dtype=DTYPE[kind], order="F")
    info = be.arnoldi(A, X, H)
    Xd = be.data(X)
    G = Xd[:, -8:].conj().T @ Xd[:, -8:]
    return {"info": info, "H": H, "x_last_head": Xd[:256, kdim].copy(), "matvecs": be.counter(A),
            "abs_orth_err": np.array(np.abs(G - np.eye(8)).max(), dtype=np.float64)}


def stencil2d_gmres(kind, be):
    """C2 (gmres row): restarted gmres on a non-symmetric 5-point convection-diffusion stencil"""
    dims = (40, 36)
    n = dims[0] * dims[1]
    A = be.stencil(kind, dims, CONVDIFF2D)
    b = be.basis_n(kind, n, 1, unit(pseudo((n,), 412, kind)))
    x = be.basis_n(kind, n, 1)
    info, meta = be.gmres(A, b, x, kdim=15, maxiter=40)
    return {"info": info, "x": be.data(x)[:, 0].copy(), "res": np.asarray(meta["res"], dtype=np.float64),
            "n_iter": meta["n_iter"], "n_inner": meta["n_inner"], "n_outer": meta["n_outer"]}


def stencil3d_lanczos(kind, be):
    """C4: lanczos on the 7-point Poisson operator"""
    dims, kdim = (12, 10, 8), 24
    n = dims[0] * dims[1] * dims[2]
    A = be.stencil(kind, dims, POISSON3D, sym=True)
    X = be.basis_n(kind, n, kdim + 1, unit(pseudo((n,), 422, kind)))
    T = np.zeros((kdim + 1, kdim), dtype=DTYPE[kind], order="F")
    info = be.lanczos(A, X, T)
    return {"info": info, "T": T, "X": be.data(X)}


def stencil3d_cg(kind, be):
    """C4 (cg row)"""
    dims = (12, 10, 8)
    n = dims[0] * dims[1] * dims[2]
    A = be.stencil(kind, dims, POISSON3D, sym=True)
    b = be.basis_n(kind, n, 1, unit(pseudo((n,), 432, kind)))
    x = be.basis_n(kind, n, 1)
    info, meta = be.cg(A, b, x, maxiter=200)
    return {"info": info, "x": be.data(x)[:, 0].copy(), "res": np.asarray(meta["res"], dtype=np.float64),
            "n_iter": meta["n_iter"]}


def stencil3d_eigs(kind, be):
    """C3: eigs (Krylov-Schur restarts) on a non-symmetric 7-point stencil"""
    dims, nev = (8, 7, 6), 2
    n = dims[0] * dims[1] * dims[2]
    A = be.stencil(kind, dims, CONVDIFF3D)
    ev, res, X, info = be.eigs(A, nev, unit(pseudo((n,), 442, kind)), kdim=20, tolerance=1e-8, n=n)
    return {"info": info, "eigvals": np.asarray(ev, dtype=np.complex128), "X": X}


def toy_csr(kind, m=72, n=56, per_row=6):
    """deterministic rectangular CSR: row i holds columns (3 i + q (n // per_row)) mod n, q < per_row (distinct), values from pseudo"""
    col = np.array([sorted((3 * i + q * (n // per_row)) % n for q in range(per_row)) for i in range(m)], dtype=np.int32).ravel()
    rowptr = np.arange(0, (m + 1) * per_row, per_row, dtype=np.int64)
    val = pseudo((m * per_row,), 601, kind)
    return m, n, rowptr, col, val


def csr_bidiag(kind, be):
    """C5: bidiagonalization of a rectangular complex CSR operator (72 x 56 here; 50M x 40M in bench_configs)"""
    m, n, rowptr, col, val = toy_csr(kind)
    kdim = 12
    A = be.csr(kind, m, n, rowptr, col, val)
    U = be.basis_n(kind, m, kdim + 1, unit(pseudo((m,), 602, kind)))
    V = be.basis_n(kind, n, kdim + 1)
    B = np.zeros((kdim + 1, kdim), dtype=DTYPE[kind], order="F")
    info = be.bidiag(A, U, V, B)
    return {"info": info, "B": B, "U": be.data(U), "V": be.data(V)}


def csr_svds(kind, be):
    """C5 (svds row)"""
    m, n, rowptr, col, val = toy_csr(kind)
    nsv = 3
    A = be.csr(kind, m, n, rowptr, col, val)
    S, res, U, V, info = be.svds(A, nsv, unit(pseudo((m,), 612, kind)), kdim=40, tolerance=1e-9, shape=(m, n))
    return {"info": info, "S": np.asarray(S, dtype=np.float64), "absU": np.abs(U), "absV": np.abs(V)}


CSR_CASES = {"csr_bidiag": csr_bidiag, "csr_svds": csr_svds}

STENCIL_CASES = {"stencil2d_arnoldi": stencil2d_arnoldi, "stencil2d_arnoldi_large": stencil2d_arnoldi_large, "stencil2d_gmres": stencil2d_gmres,
                 "stencil3d_lanczos": stencil3d_lanczos, "stencil3d_cg": stencil3d_cg, "stencil3d_eigs": stencil3d_eigs}

SOLVER_CASES = {"eighs_write_intermediate": eighs_write_intermediate, "svds_write_intermediate": svds_write_intermediate,
                "fgmres_solve": fgmres_solve, "kexpm_block": kexpm_block, "kexpm_breakdown": kexpm_breakdown, "eigs_solve": eigs_solve, "gmres_solve": gmres_solve, "cg_solve": cg_solve, "eighs_solve": eighs_solve, "svds_solve": svds_solve,
                "kexpm_solve": kexpm_solve}
SOLVER_CASES.update(STENCIL_CASES)
SOLVER_CASES.update(CSR_CASES)


def applies(name, kind):
    if name in CSR_CASES:
        return kind == "z"
    return kind == "d" if name in STENCIL_CASES else True


# ------------------------------------------------------------------------------------------------------------------ backends
SUFFIX_OF = {"s": "rsp", "d": "rdp", "c": "csp", "z": "cdp"}


class RefBackend:
    """the reference's own Fortran sources under oracle/f90run.py (container only: needs /root/reference)"""
    name = "reference"

    def __init__(self):
        from oracle import ref_exec
        self.rx = ref_exec
        assert ref_exec.test_size() == N

    def linop(self, kind, A, sym=False):
        return self.rx.linop(kind, A, sym)

    def basis(self, kind, ncols, first=None):
        return self.rx.basis(kind, ncols, first)

    def data(self, X):
        return self.rx.basis_data(X)

    def stencil(self, kind, dims, coef, sym=False):
        """user-side operator type of tests/golden/user_stencil.f90 (real(dp))"""
        it = self.rx.interp()
        if "stencil_apply_rdp" not in it.p.procs:
            import os
            from oracle import f90run
            it.p.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "user_stencil.f90"))
            f90run.Interp(it.p)
        op = it.new_inst("stencil_sym_linop_rdp" if sym else "stencil_linop_rdp")
        d = list(dims) + [1] * (3 - len(dims))
        op.f["nx"], op.f["ny"], op.f["nz"] = d
        op.f["coef"][:len(coef)] = coef
        return op

    def csr(self, kind, m, n, rowptr, col, val):
        """user-side operator type of tests/golden/user_csr.f90 (complex(dp))"""
        it = self.rx.interp()
        if "csr_matvec_cdp" not in it.p.procs:
            import os
            from oracle import f90run
            it.p.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "user_csr.f90"))
            f90run.Interp(it.p)
        op = it.new_inst("csr_linop_cdp")
        op.f["m"], op.f["n"] = m, n
        op.f["rowptr"], op.f["col"] = np.array(rowptr, dtype=np.int64), np.array(col, dtype=np.int64)
        op.f["val"] = np.array(val, dtype=np.complex128)
        return op

    def basis_n(self, kind, n, ncols, first=None):
        """the reference's own dense_vector_rdp (AbstractVectors.fypp:420-470), any length"""
        X = np.empty(ncols, dtype=object)
        for i in range(ncols):
            X[i] = self.rx.call("dense_vector", np.zeros(n, dtype=DTYPE[kind]))[0]
        if first is not None:
            first = np.asarray(first).reshape(n, -1)
            for i in range(first.shape[1]):
                X[i].f["data"][...] = first[:, i]
        return X

    def counter(self, A):
        return int(A.f["matvec_counter"])

    @staticmethod
    def _opt(**kw):
        return {k: v for k, v in kw.items() if v is not None}

    def arnoldi(self, A, X, H, kstart=None, kend=None, tol=None, transpose=None, blksize=None):
        if tol is not None:
            tol = H.real.dtype.type(tol)
        _, o = self.rx.call("arnoldi", A, X, H, 0, **self._opt(kstart=kstart, kend=kend, tol=tol, transpose=transpose,
                                                                 blksize=blksize))
        return int(o[3])

    def lanczos(self, A, X, T, kstart=None, kend=None):
        _, o = self.rx.call("lanczos", A, X, T, 0, **self._opt(kstart=kstart, kend=kend))
        return int(o[3])

    def bidiag(self, A, U, V, B, kstart=None, kend=None):
        _, o = self.rx.call("bidiagonalization", A, U, V, B, 0, **self._opt(kstart=kstart, kend=kend))
        return int(o[4])

    def qr(self, Q, tol=None):
        dt = Q[0].f["data"].dtype
        R = np.zeros((len(Q), len(Q)), dtype=dt, order="F")
        kw = {} if tol is None else {"tol": R.real.dtype.type(tol)}
        _, o = self.rx.call("qr", Q, R, 0, **kw)
        return int(o[2]), R

    def qr_pivoting(self, Q):
        dt = Q[0].f["data"].dtype
        R = np.zeros((len(Q), len(Q)), dtype=dt, order="F")
        perm = np.zeros(len(Q), dtype=np.int64)
        _, o = self.rx.call("qr", Q, R, perm, 0)
        return int(o[3]), R, perm - 1                  # 0-based like the oracle

    def krylov_schur(self, X, H):
        kind = self._kind(X)
        sel = ("procref", f"eigs_{SUFFIX_OF[kind]}::median_selector")       # the selector eigs passes (an internal procedure)
        _, o = self.rx.call("krylov_schur", 0, X, H, sel)
        return int(o[0])

    def helpers(self, kind, X, Y, Bm):
        call = self.rx.call
        one, half = DTYPE[kind](1.0), DTYPE[kind](0.5)
        out = {"innerprod_vec": np.array(call("innerprod", X, Y[0])[0]), "innerprod_mat": np.array(call("innerprod", X, Y)[0]),
               "gram": np.array(call("gram", X)[0])}
        _, o = call("linear_combination", None, X, np.ascontiguousarray(Bm[:, 0]))
        out["lincomb_vec"] = o[0].f["data"].copy()
        _, o = call("linear_combination", None, X, Bm)
        Z = o[0]
        out["lincomb_mat"] = self.data(Z)
        call("axpby_basis", half, Y, one, Z)                       # Z = 0.5 * Y + Z   (axpby_basis(alpha, X, beta, Y): Y = alpha X + beta Y)
        out["axpby_basis"] = self.data(Z)
        call("copy", Z, Y)
        out["copy"] = self.data(Z)
        call("zero_basis", Z)
        out["zero_basis"] = self.data(Z)
        return out

    def dgs_vec(self, y, X):
        dt = X[0].f["data"].dtype
        beta = np.zeros(len(X), dtype=dt)
        _, o = self.rx.call("double_gram_schmidt_step", y[0], X, 0, if_chk_orthonormal=False, beta=beta)
        return int(o[2]), beta

    def dgs_bas(self, Y, X):
        dt = X[0].f["data"].dtype
        beta = np.zeros((len(X), len(Y)), dtype=dt, order="F")
        _, o = self.rx.call("double_gram_schmidt_step", Y, X, 0, if_chk_orthonormal=False, beta=beta)
        return int(o[2]), beta


    # ---- solvers
    def _meta(self, name, kind):
        it = self.rx.interp()
        return it.new_inst(f"{name}_{'sp' if kind in 'sc' else 'dp'}_metadata")

    def _kind(self, X):
        return {np.dtype(np.float32): "s", np.dtype(np.float64): "d", np.dtype(np.complex64): "c",
                np.dtype(np.complex128): "z"}[X[0].f["data"].dtype]

    def gmres(self, A, b, x, kdim, maxiter, flexible=False):
        kind = self._kind(b)
        it = self.rx.interp()
        name = "fgmres" if flexible else "gmres"
        opts = it.new_inst(f"{name}_{'sp' if kind in 'sc' else 'dp'}_opts")
        opts.f["kdim"], opts.f["maxiter"] = kdim, maxiter
        meta = self._meta(name, kind)
        _, o = self.rx.call(name, A, b[0], x[0], 0, options=opts, meta=meta)
        m = meta.f
        return int(o[3]), {"res": np.array(m["res"][:m["n_iter"] + 1]), "n_iter": int(m["n_iter"]), "n_inner": int(m["n_inner"]),
                           "n_outer": int(m["n_outer"]), "converged": bool(m["converged"])}

    def cg(self, A, b, x, maxiter):
        kind = self._kind(b)
        it = self.rx.interp()
        opts = it.new_inst(f"cg_{'sp' if kind in 'sc' else 'dp'}_opts")
        opts.f["maxiter"] = maxiter
        meta = self._meta("cg", kind)
        _, o = self.rx.call("cg", A, b[0], x[0], 0, options=opts, meta=meta)
        m = meta.f
        return int(o[3]), {"res": np.array(m["res"][:m["n_iter"] + 1]), "n_iter": int(m["n_iter"])}

    def eighs(self, A, nev, x0, kdim, tolerance, write_intermediate=False):
        kind = {np.dtype(np.float32): "s", np.dtype(np.float64): "d", np.dtype(np.complex64): "c",
                np.dtype(np.complex128): "z"}[x0.dtype]
        X = self.basis(kind, nev)
        x0v = self.rx.vector(kind, x0)
        _, o = self.rx.call("eighs", A, X, None, None, 0, x0=x0v, kdim=kdim, tolerance=_real(kind)(tolerance),
                            write_intermediate=write_intermediate)
        return np.array(o[2]), np.array(o[3]), self.data(X), int(o[4])

    def eigs(self, A, nev, x0, kdim, tolerance, n=None):
        kind = {np.dtype(np.float32): "s", np.dtype(np.float64): "d", np.dtype(np.complex64): "c",
                np.dtype(np.complex128): "z"}[x0.dtype]
        if n is None:
            X = self.basis(kind, nev)
            x0v = self.rx.vector(kind, x0)
        else:
            X = self.basis_n(kind, n, nev)
            x0v = self.basis_n(kind, n, 1, x0)[0]
        _, o = self.rx.call("eigs", A, X, None, None, 0, x0=x0v, kdim=kdim, tolerance=_real(kind)(tolerance),
                            write_intermediate=False)
        return np.array(o[2]), np.array(o[3]), self.data(X), int(o[4])

    def svds(self, A, nsv, u0, kdim, tolerance, write_intermediate=False, shape=None):
        kind = {np.dtype(np.float32): "s", np.dtype(np.float64): "d", np.dtype(np.complex64): "c",
                np.dtype(np.complex128): "z"}[u0.dtype]
        if shape is None:
            U, V = self.basis(kind, nsv), self.basis(kind, nsv)
            u0v = self.rx.vector(kind, u0)
        else:
            U, V = self.basis_n(kind, shape[0], nsv), self.basis_n(kind, shape[1], nsv)
            u0v = self.basis_n(kind, shape[0], 1, u0)[0]
        _, o = self.rx.call("svds", A, U, None, V, None, 0, u0=u0v, kdim=kdim, tolerance=_real(kind)(tolerance),
                            write_intermediate=write_intermediate)
        return np.array(o[2]), np.array(o[4]), self.data(U), self.data(V), int(o[5])

    def kexpm(self, c, A, b, tau, tol, kdim):
        kind = self._kind(b)
        rt = _real(kind)
        _, o = self.rx.call("kexpm", c[0], A, b[0], rt(tau), rt(tol), 0, kdim=kdim)
        return int(o[5])

    def kexpm_mat(self, Cb, A, B, tau, tol, kdim):
        rt = _real(self._kind(B))
        _, o = self.rx.call("kexpm", Cb, A, B, rt(tau), rt(tol), 0, kdim=kdim)
        return int(o[5])


class OracleBackend:
    """oracle/lk_oracle (C restatement) -- the thing the fixtures pin"""
    name = "oracle"

    def __init__(self):
        from oracle import lk_oracle
        self.lo = lk_oracle

    def linop(self, kind, A, sym=False):
        return self.lo.Op.dense(np.asfortranarray(A))

    def basis(self, kind, ncols, first=None):
        X = np.zeros((N, ncols), dtype=DTYPE[kind], order="F")
        if first is not None:
            first = np.asarray(first)
            if first.ndim == 1:
                first = first[:, None]
            X[:, :first.shape[1]] = first
        return X

    def data(self, X):
        return X

    def stencil(self, kind, dims, coef, sym=False):
        return self.lo.Op.stencil(kind, tuple(dims), tuple(float(c) for c in coef))

    def csr(self, kind, m, n, rowptr, col, val):
        return self.lo.Op.csr(m, n, rowptr, col, val)

    def basis_n(self, kind, n, ncols, first=None):
        X = np.zeros((n, ncols), dtype=DTYPE[kind], order="F")
        if first is not None:
            first = np.asarray(first).reshape(n, -1)
            X[:, :first.shape[1]] = first
        return X

    def counter(self, A):
        return int(A.n_matvec)

    def arnoldi(self, A, X, H, kstart=None, kend=None, tol=None, transpose=None, blksize=None):
        return self.lo.arnoldi(A, X, H, kstart=kstart or 1, kend=kend, tol=tol, trans=bool(transpose), blksize=blksize or 1)

    def lanczos(self, A, X, T, kstart=None, kend=None):
        return self.lo.lanczos(A, X, T, kstart=kstart or 1, kend=kend)

    def bidiag(self, A, U, V, B, kstart=None, kend=None):
        return self.lo.bidiag(A, U, V, B, kstart=kstart or 1, kend=kend)

    def qr(self, Q, tol=None):
        return self.lo.qr(Q, tol=tol)

    def qr_pivoting(self, Q):
        return self.lo.qr_with_pivoting(Q)

    def krylov_schur(self, X, H):
        return self.lo.krylov_schur(X, H)

    def helpers(self, kind, X, Y, Bm):
        """the same operation order as the reference's loops, on the oracle's dot / axpby primitives"""
        lo, dt = self.lo, DTYPE[kind]
        j, p = X.shape[1], Y.shape[1]
        col = np.ascontiguousarray
        ip = np.array([[lo.dot(col(X[:, i]), col(Y[:, q])) for q in range(p)] for i in range(j)], dtype=dt).reshape(j, p)
        # LITERAL: gram_matrix_* fills the lower triangle with G(j, i) = G(i, j) -- NOT conjugated, also for the complex kinds
        # (AbstractVectors.fypp:645-660), so the reference's complex Gram matrix is symmetric, not Hermitian.  Harmless where the
        # reference uses it (is_orthonormal takes the Frobenius norm of G - I); device vectors inherit it, the reference's own
        # Gram runs on their type-bound dot.
        gram = np.zeros((j, j), dtype=dt)
        for i in range(j):
            for q in range(i, j):
                gram[i, q] = lo.dot(col(X[:, i]), col(X[:, q]))
                gram[q, i] = gram[i, q]
        Z = np.zeros((N, p), dtype=dt, order="F")
        for q in range(p):
            z = np.zeros(N, dtype=dt)
            for i in range(j):
                lo.axpby(Bm[i, q], col(X[:, i]), 1.0, z)
            Z[:, q] = z
        out = {"innerprod_vec": ip[:, 0].copy(), "innerprod_mat": ip, "gram": gram, "lincomb_vec": Z[:, 0].copy(),
               "lincomb_mat": Z.copy(order="F")}
        Z2 = (dt(0.5) * Y + Z).astype(dt)
        out["axpby_basis"] = Z2
        out["copy"] = Y.copy(order="F")
        out["zero_basis"] = np.zeros_like(Y)
        return out

    def dgs_vec(self, y, X):
        return self.lo.dgs_vec(y[:, 0], X, X.shape[1])

    def dgs_bas(self, Y, X):
        return self.lo.dgs_bas(Y, X, X.shape[1])

    # ---- solvers
    def gmres(self, A, b, x, kdim, maxiter, flexible=False):
        info, meta = self.lo.gmres(A, b[:, 0], x[:, 0], kdim=kdim, maxiter=maxiter, flexible=flexible)
        return info, meta

    def cg(self, A, b, x, maxiter):
        info, meta = self.lo.cg(A, b[:, 0], x[:, 0], maxiter=maxiter)
        return info, meta

    def eighs(self, A, nev, x0, kdim, tolerance, write_intermediate=False):
        ev, res, X, k = self.lo.eighs(A, N, nev, x0, kdim=kdim, tolerance=tolerance, write_intermediate=write_intermediate)
        return ev, res, X, k

    def eigs(self, A, nev, x0, kdim, tolerance, n=None):
        ev, res, X, niter = self.lo.eigs(A, N if n is None else n, nev, x0, kdim=kdim, tolerance=tolerance)
        return ev, res, X, niter

    def svds(self, A, nsv, u0, kdim, tolerance, write_intermediate=False, shape=None):
        S, res, U, V, k = self.lo.svds(A, nsv, u0, kdim=kdim, tolerance=tolerance, write_intermediate=write_intermediate)
        return S, res, U, V, k

    def kexpm(self, c, A, b, tau, tol, kdim):
        out, info = self.lo.kexpm_vec(A, b[:, 0].copy(), tau, tol, kdim=kdim)
        c[:, 0] = out
        return info

    def kexpm_mat(self, Cb, A, B, tau, tol, kdim):
        out, info = self.lo.kexpm_mat(A, np.asfortranarray(B.copy()), tau, tol, kdim=kdim)
        Cb[...] = out
        return info
