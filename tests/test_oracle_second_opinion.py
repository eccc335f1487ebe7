"""A SECOND, independent restatement of the reference's step loops, used only to cross-check the C oracle (oracle/).

The C oracle (oracle/lk_oracle.c + lko_body.inc) is what every GPU parity test compares against, and nothing in this
environment can COMPILE the Fortran reference (no compiler here or on the GPU box; since the end of round 2 its sources are
executed by the interpreter oracle/f90run.py instead -- tests/test_ref_golden.py, DESIGN.md section 4 -- this file predates
that and stays as a further, independent check).  To reduce the risk that the oracle and the product share one misreading of the
Fortran, this file restates the same routines a second time -- in plain numpy, written directly from the reference sources
cited below, sharing NO code with oracle/ -- and requires the two restatements to agree entry by entry:

    arnoldi                     /root/reference/src/Krylov/arnoldi.fypp:34-73          (blksize = 1 and > 1)
    double_gram_schmidt_step    src/Krylov/gram_schmidt.fypp:12-57, 59-105  = two passes of
    orthogonalize_against_basis src/Krylov/gram_schmidt.fypp:113-154, 156-200  (innerprod, linear_combination, sub)
    qr_no_pivoting              src/Krylov/qr.fypp:116-167
    lanczos + update_tridiag    src/Krylov/lanczos.fypp:7-64
    bidiagonalization           src/Krylov/golub_kahan.fypp:7-64
    vector semantics            src/Utilities/TestUtils.fypp:240-300 (dot_product conjugates `self`; axpby: self = alpha*vec + beta*self)
                                src/AbstractTypes/AbstractVectors.fypp:424-460 (norm = sqrt(abs(dot(x, x))), sub, chsgn), :571-695

CPU only (no GPU, no product code).  Tolerance: the two restatements use different summation orders (numpy pairwise / BLAS
vs the oracle's sequential OpenMP loops), so entries agree to rounding: 1e-12 in fp64, 2e-4 in fp32 (normwise).
"""
import numpy as np
import pytest

from helpers import rel_normwise, randn

ATOL = {"s": 1e-6, "d": 1e-15, "c": 1e-6, "z": 1e-15}            # src/Constants.f90:16-48
DT = {"s": np.float32, "d": np.float64, "c": np.complex64, "z": np.complex128}


# ---- abstract_vector semantics (each vector is a column of a numpy array, updated in place) ---------------------------
def v_dot(x, y):                       # self%dot(vec) = dot_product(self, vec): conjugate on self
    return np.vdot(x, y).astype(x.dtype)


def v_norm(x):                         # AbstractVectors.fypp:424-432
    return np.sqrt(np.abs(v_dot(x, x))).real.astype(x.real.dtype)


def v_axpby(alpha, x, beta, y):        # self (= y) <- alpha * vec (= x) + beta * self
    y[:] = (np.asarray(alpha, dtype=y.dtype) * x + np.asarray(beta, dtype=y.dtype) * y).astype(y.dtype)


def innerprod(X, y):                   # AbstractVectors.fypp:659-675: v(i) = X(i)%dot(y)
    return np.array([v_dot(X[:, i], y) for i in range(X.shape[1])], dtype=y.dtype)


def linear_combination(X, v):          # :571-603: y = 0 ; y%axpby(v(i), X(i), 1) for every i
    y = np.zeros(X.shape[0], dtype=X.dtype)
    for i in range(X.shape[1]):
        v_axpby(v[i], X[:, i], 1, y)
    return y


def orthogonalize_vector_against_basis(y, X, kind):      # gram_schmidt.fypp:113-154 (if_chk_orthonormal = .false.)
    info = 1 if v_norm(y) < ATOL[kind] else 0
    c = innerprod(X, y)
    proj = linear_combination(X, c)
    v_axpby(-1, proj, 1, y)                                # y%sub(proj)
    return c, info


def dgs_vector(y, X, kind):                              # gram_schmidt.fypp:12-57
    c1, info = orthogonalize_vector_against_basis(y, X, kind)
    c2, info = orthogonalize_vector_against_basis(y, X, kind)      # `info` is overwritten by the second pass
    return (c1 + c2).astype(y.dtype), info


def dgs_basis(Y, X, kind):                               # gram_schmidt.fypp:59-105 + 156-200: per-pass, all columns of Y
    C = np.zeros((X.shape[1], Y.shape[1]), dtype=Y.dtype)
    info = 0
    for _ in range(2):
        info = 0
        for i in range(Y.shape[1]):
            if v_norm(Y[:, i]) < ATOL[kind]:
                info = i + 1
        P = np.array([[v_dot(X[:, a], Y[:, b]) for b in range(Y.shape[1])] for a in range(X.shape[1])], dtype=Y.dtype)
        for b in range(Y.shape[1]):
            proj = linear_combination(X, P[:, b])
            v_axpby(-1, proj, 1, Y[:, b])
        C += P
    return C, info


def qr_no_pivoting(Q, kind, tol=None):                   # qr.fypp:116-167 (no breakdown is provoked in these tests)
    tol = ATOL[kind] if tol is None else tol
    p = Q.shape[1]
    R = np.zeros((p, p), dtype=Q.dtype)
    info = 0
    for j in range(p):
        if j > 0:
            R[:j, j], info = dgs_vector(Q[:, j], Q[:, :j], kind)
        beta = v_norm(Q[:, j])
        assert np.isfinite(beta) and beta >= tol, "these inputs must not break down"
        R[j, j] = beta
        Q[:, j] *= Q.dtype.type(1) / Q.dtype.type(beta)
    return R, info


def arnoldi(apply_A, X, H, kind, p=1, tol=None):         # arnoldi.fypp:34-73
    tol = ATOL[kind] if tol is None else tol
    kdim = (X.shape[1] - p) // p
    for k in range(1, kdim + 1):
        kpm, kp, kpp = (k - 1) * p, k * p, (k + 1) * p
        for i in range(p):
            X[:, kp + i] = apply_A(X[:, kpm + i])
        if p == 1:
            H[:kp, kpm], _ = dgs_vector(X[:, kp], X[:, :kp], kind)
        else:
            H[:kp, kpm:kp], _ = dgs_basis(X[:, kp:kpp], X[:, :kp], kind)
        R, _ = qr_no_pivoting(X[:, kp:kpp], kind)
        H[kp:kpp, kpm:kp] = R
        if min(abs(R[i, i]) for i in range(p)) < tol:
            return kp
    return 0


def lanczos(apply_A, X, T, kind, tol=None):              # lanczos.fypp:7-64
    tol = ATOL[kind] if tol is None else tol
    kdim = X.shape[1] - 1
    for k in range(1, kdim + 1):
        X[:, k] = apply_A(X[:, k - 1])
        for i in range(max(1, k - 1), k + 1):            # update_tridiag_matrix
            T[i - 1, k - 1] = v_dot(X[:, i - 1], X[:, k])
            v_axpby(-T[i - 1, k - 1], X[:, i - 1], 1, X[:, k])
        dgs_vector(X[:, k], X[:, :k], kind)
        beta = v_norm(X[:, k])
        T[k, k - 1] = beta
        if beta < tol:
            return k
        X[:, k] *= X.dtype.type(1) / X.dtype.type(beta)
    return 0


def bidiagonalization(A, U, V, B, kind, tol=None):       # golub_kahan.fypp:7-64
    tol = ATOL[kind] if tol is None else tol
    kdim = U.shape[1] - 1
    for k in range(1, kdim + 1):
        V[:, k - 1] = (A.conj().T @ U[:, k - 1]).astype(V.dtype)
        if k > 1:
            dgs_vector(V[:, k - 1], V[:, :k - 1], kind)
        alpha = v_norm(V[:, k - 1]); B[k - 1, k - 1] = alpha
        if not abs(alpha) > tol:
            return k
        V[:, k - 1] *= V.dtype.type(1) / V.dtype.type(alpha)
        U[:, k] = (A @ V[:, k - 1]).astype(U.dtype)
        dgs_vector(U[:, k], U[:, :k], kind)
        beta = v_norm(U[:, k]); B[k, k - 1] = beta
        if not abs(beta) > tol:
            return k
        U[:, k] *= U.dtype.type(1) / U.dtype.type(beta)
    return 0


# ---- the cross-checks ---------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def oracle():
    from oracle import lk_oracle
    return lk_oracle


def _tol(kind):
    return 1e-12 if kind in "dz" else 2e-4


@pytest.mark.parametrize("kind", ["s", "d", "c", "z"])
def test_arnoldi_config1_two_restatements_agree(oracle, kind):
    """BASELINE config 1 (TestKrylov-style: random dense operator, n = 128, kdim = 64)."""
    dt = DT[kind]; n, kdim = 128, 64
    rng = np.random.default_rng(1)
    A = randn(rng, (n, n), dt)
    x0 = randn(rng, n, dt); x0 /= np.linalg.norm(x0)
    X2 = np.zeros((n, kdim + 1), dtype=dt, order="F"); X2[:, 0] = x0
    H2 = np.zeros((kdim + 1, kdim), dtype=dt, order="F")
    assert arnoldi(lambda v: (A @ v).astype(dt), X2, H2, kind) == 0
    Xo = np.zeros((n, kdim + 1), dtype=dt, order="F"); Xo[:, 0] = x0
    Ho = np.zeros((kdim + 1, kdim), dtype=dt, order="F")
    assert oracle.arnoldi(oracle.Op.dense(A), Xo, Ho) == 0
    assert rel_normwise(Ho, H2) < _tol(kind)
    assert rel_normwise(Xo, X2) < (1e-10 if kind in "dz" else 5e-3)
    # and both satisfy the reference's own assertion  A X_k = X_{k+1} H   (test/TestKrylov.fypp:194-242)
    assert np.abs(A @ X2[:, :kdim] - X2 @ H2).max() < (1e-12 if kind in "dz" else 1e-3)


@pytest.mark.parametrize("kind", ["d", "z"])
@pytest.mark.parametrize("p", [2, 3])
def test_block_arnoldi_two_restatements_agree(oracle, kind, p):
    dt = DT[kind]; n, kdim = 96, 10
    rng = np.random.default_rng(2)
    A = randn(rng, (n, n), dt)
    X0, _ = np.linalg.qr(randn(rng, (n, p), dt))
    X2 = np.zeros((n, (kdim + 1) * p), dtype=dt, order="F"); X2[:, :p] = X0
    H2 = np.zeros(((kdim + 1) * p, kdim * p), dtype=dt, order="F")
    assert arnoldi(lambda v: (A @ v).astype(dt), X2, H2, kind, p=p) == 0
    Xo = np.zeros_like(X2); Xo[:, :p] = X0
    Ho = np.zeros_like(H2)
    assert oracle.arnoldi(oracle.Op.dense(A), Xo, Ho, blksize=p) == 0
    assert rel_normwise(Ho, H2) < 1e-12


@pytest.mark.parametrize("kind", ["s", "d", "c", "z"])
def test_lanczos_two_restatements_agree(oracle, kind):
    dt = DT[kind]; n, kdim = 128, 40
    rng = np.random.default_rng(3)
    M = randn(rng, (n, n), dt)
    A = np.asfortranarray(((M + M.conj().T) / 2).astype(dt))                  # symmetric / Hermitian
    x0 = randn(rng, n, dt); x0 /= np.linalg.norm(x0)
    X2 = np.zeros((n, kdim + 1), dtype=dt, order="F"); X2[:, 0] = x0
    T2 = np.zeros((kdim + 1, kdim), dtype=dt, order="F")
    assert lanczos(lambda v: (A @ v).astype(dt), X2, T2, kind) == 0
    Xo = np.zeros_like(X2); Xo[:, 0] = x0
    To = np.zeros_like(T2)
    assert oracle.lanczos(oracle.Op.dense(A), Xo, To) == 0
    assert rel_normwise(To, T2) < _tol(kind)


@pytest.mark.parametrize("kind", ["s", "d", "c", "z"])
def test_bidiagonalization_two_restatements_agree(oracle, kind):
    dt = DT[kind]; m, n, kdim = 120, 90, 30
    rng = np.random.default_rng(4)
    A = randn(rng, (m, n), dt)
    u0 = randn(rng, m, dt); u0 /= np.linalg.norm(u0)
    U2 = np.zeros((m, kdim + 1), dtype=dt, order="F"); U2[:, 0] = u0
    V2 = np.zeros((n, kdim + 1), dtype=dt, order="F"); B2 = np.zeros((kdim + 1, kdim), dtype=dt, order="F")
    assert bidiagonalization(A, U2, V2, B2, kind) == 0
    Uo = np.zeros_like(U2); Uo[:, 0] = u0
    Vo = np.zeros_like(V2); Bo = np.zeros_like(B2)
    assert oracle.bidiag(oracle.Op.dense(A), Uo, Vo, Bo) == 0
    assert rel_normwise(Bo, B2) < _tol(kind)


# ---- solver shells: gmres (GMRES/gmres.fypp:100-238 + submodule_utility_functions.fypp:167-204) and cg (CG/CG.fypp:96-179) ----
RTOL = {k: np.sqrt(v) for k, v in ATOL.items()}                  # src/Constants.f90: rtol = sqrt(atol)


def apply_givens_rotation(h, c, s):
    """h(1:k+1) = new Hessenberg column; c, s = rotations so far, entry k is produced here."""
    k = h.size - 1
    if np.iscomplexobj(h):
        for i in range(k - 1):                                   # hand-rolled: NO conjugates (literal restatement)
            t = c[i] * h[i] + s[i] * h[i + 1]
            h[i + 1] = -s[i] * h[i] + c[i] * h[i + 1]
            h[i] = t
        nrm = np.sqrt(abs(h[k - 1]) ** 2 + abs(h[k]) ** 2)       # g = x / norm(x, 2)
        c[k - 1], s[k - 1] = h[k - 1] / nrm, h[k] / nrm
        h[k - 1] = c[k - 1] * h[k - 1] + s[k - 1] * h[k]
        h[k] = 0
    else:
        for j in range(k - 1):                                   # lasr('L', 'V', 'F'): P(k-1) ... P(1) from the left
            t = h[j + 1]
            h[j + 1] = c[j] * t - s[j] * h[j]
            h[j] = s[j] * t + c[j] * h[j]
        f, g = h[k - 1], h[k]                                    # lartg (LAPACK 3.10: r = sign(f) sqrt(f^2 + g^2), c >= 0)
        if g == 0:
            c[k - 1], s[k - 1], r = 1.0, 0.0, f
        elif f == 0:
            c[k - 1], s[k - 1], r = 0.0, np.sign(g), abs(g)
        else:
            d = np.sqrt(f * f + g * g)
            c[k - 1] = abs(f) / d
            r = np.copysign(d, f)
            s[k - 1] = g / r
        h[k - 1] = r
        h[k] = 0


def gmres(apply_A, b, x, kind, kdim=30, maxiter=10, precond=None, flexible=False):
    """gmres.fypp:125-238; flexible = fgmres.fypp:130-215 (Z(k) = M_k^-1 V(k) stored, dx = Z(:k) y, no post-application)."""
    dt = b.dtype
    tol = ATOL[kind] + RTOL[kind] * v_norm(b)
    n = b.size
    res, n_iter, n_inner, n_outer, converged = [], 0, 0, 0, False
    while (not converged) and n_outer <= maxiter:
        H = np.zeros((kdim + 1, kdim), dtype=dt); V = np.zeros((n, kdim + 1), dtype=dt, order="F")
        if v_norm(x) != 0:
            V[:, 0] = apply_A(x)
        v_axpby(-1, b, 1, V[:, 0]); V[:, 0] *= -1                # sub(b) ; chsgn()
        e = np.zeros(kdim + 1, dtype=dt)
        beta = v_norm(V[:, 0]); e[0] = beta
        V[:, 0] *= dt.type(1) / dt.type(beta)
        c = np.zeros(kdim, dtype=dt); s = np.zeros(kdim, dtype=dt)
        if n_outer == 0:
            res.append(abs(beta))
        k = 0
        Z = np.zeros((n, kdim), dtype=dt, order="F")
        for k in range(1, kdim + 1):
            wrk = V[:, k - 1].copy()                               # wrk = V(k)   /   copy(Z(k), V(k))
            if precond is not None:
                wrk = precond(wrk, k)                              # preconditioner%apply(wrk, k, beta, tol)
            Z[:, k - 1] = wrk
            V[:, k] = apply_A(wrk)
            H[:k, k - 1], _ = dgs_vector(V[:, k], V[:, :k], kind)
            H[k, k - 1] = v_norm(V[:, k])
            if abs(H[k, k - 1]) > tol:
                V[:, k] *= dt.type(1) / H[k, k - 1]
            apply_givens_rotation(H[:k + 1, k - 1], c[:k], s[:k])
            e[k] = -s[k - 1] * e[k - 1]; e[k - 1] = c[k - 1] * e[k - 1]
            beta = abs(e[k])
            n_iter += 1; n_inner += 1; res.append(abs(beta))
            if abs(beta) < tol:
                converged = True
                break
        else:
            k = kdim + 1                                         # Fortran: the loop variable ends at kdim + 1
        k = min(k, kdim)
        y = np.linalg.solve(np.triu(H[:k, :k]), e[:k])           # trtrs('u', 'n', 'n')
        dx = linear_combination((Z if flexible else V)[:, :k], y.astype(dt))
        if precond is not None and not flexible:
            dx = precond(dx, 0)                                    # preconditioner%apply(dx)
        v_axpby(1, dx, 1, x)
        V[:, 0] = apply_A(x); v_axpby(-1, b, 1, V[:, 0]); V[:, 0] *= -1
        beta = v_norm(V[:, 0])
        n_iter += 1; n_outer += 1; res.append(abs(beta))
        if abs(beta) < tol:
            converged = True
            break
    return (n_iter if converged else -n_iter), dict(n_iter=n_iter, n_inner=n_inner, n_outer=n_outer, res=res, converged=converged)


def cg(apply_A, b, x, kind, maxiter=100):
    dt = b.dtype
    tol = ATOL[kind] + RTOL[kind] * v_norm(b)
    r = np.zeros_like(b)
    if v_norm(x) > 0:
        r = apply_A(x)
    v_axpby(-1, b, 1, r); r *= -1
    p = r.copy(); rr_old = v_dot(r, r)
    res, n_iter, converged = [np.sqrt(abs(rr_old))], 0, False
    for _ in range(maxiter):
        Ap = apply_A(p)
        alpha = rr_old / v_dot(p, Ap)
        v_axpby(alpha, p, 1, x)
        v_axpby(-alpha, Ap, 1, r)
        rr_new = v_dot(r, r)
        residual = np.sqrt(abs(rr_new))
        n_iter += 1; res.append(residual)
        if residual < tol:
            converged = True
            break
        beta = rr_new / rr_old
        v_axpby(1, r, beta, p)
        rr_old = rr_new
    return (n_iter if converged else -n_iter), dict(n_iter=n_iter, res=res, converged=converged)


@pytest.mark.parametrize("kind", ["d", "z"])
def test_gmres_two_restatements_agree(oracle, kind):
    """Restarted GMRES with several outer cycles: identical info / iteration counters, residual history and solution."""
    dt = DT[kind]; n = 200
    rng = np.random.default_rng(5)
    A = np.asfortranarray((np.eye(n) * 6 + randn(rng, (n, n), dt) * 0.25).astype(dt))
    b = randn(rng, n, dt)
    x2 = np.zeros(n, dtype=dt); xo = np.zeros(n, dtype=dt)
    # kind "z": the reference's hand-rolled complex rotation (submodule_utility_functions.fypp:173-190) uses c and s WITHOUT
    # conjugates, so it is not unitary and |e(k+1)| is not the true residual: on this non-Hermitian matrix the restarted
    # iteration stalls and later diverges (rounding differences between two evaluations are then amplified with it).
    # Both restatements reproduce that literally; they are compared over the first three cycles, entry by entry relative to
    # each entry.  (The product is tested against the same literal behaviour: test_gmres_vs_oracle[z].)
    maxiter = 30 if kind == "d" else 2
    info2, m2 = gmres(lambda v: (A @ v).astype(dt), b, x2, kind, kdim=12, maxiter=maxiter)
    infoo, mo = oracle.gmres(oracle.Op.dense(A), b, xo, kdim=12, maxiter=maxiter)
    assert info2 == infoo and m2["n_outer"] == mo["n_outer"] >= 2 and m2["n_inner"] == mo["n_inner"]
    assert len(m2["res"]) == len(mo["res"])
    r2, ro = np.array(m2["res"]), np.array(mo["res"])
    assert (np.abs(r2 - ro) / np.abs(ro)).max() < 1e-8
    assert np.linalg.norm(x2 - xo) < 1e-8 * np.linalg.norm(xo)
    if kind == "d":
        assert info2 > 0 and np.linalg.norm(A @ x2 - b) < 1e-6 * np.linalg.norm(b)
    else:
        # literal behaviour: the recomputed residual at the end of every cycle (entries 13, 26, 39) is LARGER than the last
        # inner estimate |e(k+1)| -- the rotation is not norm-preserving
        assert info2 < 0 and ro[13] > ro[12] and ro[26] > ro[25] and ro[39] > ro[38]


@pytest.mark.parametrize("kind", ["d", "z"])
def test_cg_two_restatements_agree(oracle, kind):
    dt = DT[kind]; n = 160
    rng = np.random.default_rng(6)
    M = randn(rng, (n, n), dt)
    A = np.asfortranarray((M @ M.conj().T / n + np.eye(n)).astype(dt))         # SPD / HPD
    b = randn(rng, n, dt)
    x2 = np.zeros(n, dtype=dt); xo = np.zeros(n, dtype=dt)
    info2, m2 = cg(lambda v: (A @ v).astype(dt), b, x2, kind, maxiter=200)
    infoo, mo = oracle.cg(oracle.Op.dense(A), b, xo, maxiter=200)
    assert info2 == infoo > 0
    assert np.abs(np.array(m2["res"]) - np.array(mo["res"])).max() < 1e-11 * mo["res"][0]
    assert np.linalg.norm(x2 - xo) < 1e-11 * np.linalg.norm(xo)


# ---- eighs (EIGHS/eighs.fypp:60-125) and svds (SVDS/svd_solvers.fypp:60-121): one Krylov step + one small dense factorisation
# per iteration, residual = |beta * last row of the Ritz vectors| (IterativeSolvers.fypp:930-939), stop when nev converged ----
def eighs(apply_A, n, nev, x0, kind, kdim, tol):
    dt = DT[kind]
    Xw = np.zeros((n, kdim + 1), dtype=dt, order="F"); Xw[:, 0] = x0 / v_norm(x0)
    T = np.zeros((kdim + 1, kdim), dtype=dt)
    ev = np.zeros(kdim); res = np.zeros(kdim); vec = np.zeros((kdim, kdim), dtype=dt)
    k = 0
    for k in range(1, kdim + 1):
        # lanczos(A, Xwrk, T, info, kstart = k, kend = k): ONE step of the loop above
        Xw[:, k] = apply_A(Xw[:, k - 1])
        for i in range(max(1, k - 1), k + 1):
            T[i - 1, k - 1] = v_dot(Xw[:, i - 1], Xw[:, k])
            v_axpby(-T[i - 1, k - 1], Xw[:, i - 1], 1, Xw[:, k])
        dgs_vector(Xw[:, k], Xw[:, :k], kind)
        beta = v_norm(Xw[:, k]); T[k, k - 1] = beta
        Xw[:, k] *= dt(1) / dt(beta)
        ev[:] = 0; vec[:] = 0
        Tk = T[:k, :k]
        w, v = np.linalg.eigh(np.tril(Tk) + np.tril(Tk, -1).conj().T)          # eigh reads one triangle (lower here, as syev 'L')
        ev[:k] = w; vec[:k, :k] = v
        res[:k] = np.abs(beta * vec[k - 1, :k])
        if int((res[:k] < tol).sum()) >= nev:
            break
    idx = np.argsort(-ev, kind="stable")                                        # sort_index(reverse = .true.): non-increasing, ties in original order
    k = min(k, kdim)
    return ev[idx[:nev]], res[idx[:nev]], k


def svds(A, nsv, u0, kind, kdim, tol):
    dt = DT[kind]; m, n = A.shape
    Uw = np.zeros((m, kdim + 1), dtype=dt, order="F"); Uw[:, 0] = u0 / v_norm(u0)
    Vw = np.zeros((n, kdim + 1), dtype=dt, order="F"); B = np.zeros((kdim + 1, kdim), dtype=dt)
    sv = np.zeros(kdim); res = np.zeros(kdim)
    k = 0
    for k in range(1, kdim + 1):
        # bidiagonalization(A, Uwrk, Vwrk, B, info, kstart = k, kend = k, tol = tol): ONE step
        Vw[:, k - 1] = (A.conj().T @ Uw[:, k - 1]).astype(dt)
        if k > 1:
            dgs_vector(Vw[:, k - 1], Vw[:, :k - 1], kind)
        alpha = v_norm(Vw[:, k - 1]); B[k - 1, k - 1] = alpha
        assert abs(alpha) > tol
        Vw[:, k - 1] *= dt(1) / dt(alpha)
        Uw[:, k] = (A @ Vw[:, k - 1]).astype(dt)
        dgs_vector(Uw[:, k], Uw[:, :k], kind)
        beta = v_norm(Uw[:, k]); B[k, k - 1] = beta
        assert abs(beta) > tol
        Uw[:, k] *= dt(1) / dt(beta)
        u, s, vh = np.linalg.svd(B[:k, :k])
        vmat = vh.conj().T                                                      # vmat = hermitian(vmat)
        sv[:] = 0; sv[:k] = s
        res[:k] = np.abs(B[k, k - 1] * vmat[k - 1, :k])
        if int((res[:k] < tol).sum()) >= nsv:
            break
    k = min(k, kdim)
    return sv[:nsv].copy(), res[:nsv].copy(), k


@pytest.mark.parametrize("kind", ["d", "z"])
def test_eighs_two_restatements_agree(oracle, kind):
    dt = DT[kind]; n, nev, kdim = 150, 4, 60
    rng = np.random.default_rng(7)
    Q, _ = np.linalg.qr(randn(rng, (n, n), dt))
    lam = np.concatenate([[40.0, 33.0, 27.0, 22.0], np.linspace(0.0, 10.0, n - 4)])
    A = np.asfortranarray(((Q * lam) @ Q.conj().T).astype(dt)); A = np.asfortranarray(((A + A.conj().T) / 2).astype(dt))
    x0 = randn(rng, n, dt)
    ev2, res2, k2 = eighs(lambda v: (A @ v).astype(dt), n, nev, x0.copy(), kind, kdim, 1e-8)
    evo, reso, _, ko = oracle.eighs(oracle.Op.dense(A), n, nev, x0.copy(), kdim=kdim, tolerance=1e-8)
    assert k2 == ko < kdim                                      # same iteration count (= info), converged before kdim
    assert np.abs(ev2 - evo).max() < 1e-11 * np.abs(evo).max()
    assert np.abs(ev2 - lam[:4]).max() < 1e-7                   # and the known answer
    assert np.abs(res2 - reso).max() < 1e-9


@pytest.mark.parametrize("kind", ["d", "z"])
def test_svds_two_restatements_agree(oracle, kind):
    dt = DT[kind]; m, n, nsv, kdim = 140, 100, 3, 60
    rng = np.random.default_rng(8)
    U, _ = np.linalg.qr(randn(rng, (m, n), dt)); V, _ = np.linalg.qr(randn(rng, (n, n), dt))
    sig = np.concatenate([[30.0, 24.0, 19.0], np.linspace(8.0, 0.5, n - 3)])
    A = np.asfortranarray(((U * sig) @ V.conj().T).astype(dt))
    u0 = randn(rng, m, dt)
    s2, res2, k2 = svds(A, nsv, u0.copy(), kind, kdim, 1e-8)
    so, reso, _, _, ko = oracle.svds(oracle.Op.dense(A), nsv, u0.copy(), kdim=kdim, tolerance=1e-8)
    assert k2 == ko < kdim
    assert np.abs(s2 - so).max() < 1e-11 * so.max() and np.abs(s2 - sig[:3]).max() < 1e-7
    assert np.abs(res2 - reso).max() < 1e-9


# ---- eigs with Krylov-Schur restarts, LITERAL control flow of IterativeSolvers.fypp:1052-1131 + BaseKrylov.fypp:782-834 --------
def _eig_real_pair(Hk):
    """stdlib `eig` on a real matrix = geev: complex eigenvalues, eigenvectors in LAPACK's real-pair storage."""
    from scipy.linalg import lapack
    wr, wi, vl, vr, info = lapack.dgeev(np.asfortranarray(Hk), compute_vl=0, compute_vr=1)
    assert info == 0
    return wr + 1j * wi, vr


def _krylov_schur_literal(X, H):
    from scipy.linalg import lapack
    kdim = X.shape[1] - 1
    cplx = np.iscomplexobj(H)
    Hk = np.asfortranarray(H[:kdim, :kdim])
    if cplx:
        T, sdim, w, Z, work, info = lapack.zgees(lambda x: False, Hk, sort_t=0); ev = w
    else:
        T, sdim, wr, wi, Z, work, info = lapack.dgees(lambda x, y: False, Hk, sort_t=0); ev = wr + 1j * wi
    assert info == 0
    selected = np.abs(ev) > np.median(np.abs(ev))                 # median_selector, IterativeSolvers.fypp:1136-1141
    n = int(selected.sum())
    out = (lapack.ztrsen if cplx else lapack.dtrsen)(selected.astype(np.int32), T, Z, job="N", wantq=1)
    T, Z = out[0], out[1]
    assert out[-1] == 0
    Xn = X[:, :kdim] @ Z[:, :n]                                    # linear_combination(Xwrk, X(:kdim), Z(:, :n))
    b = H[kdim, :] @ Z
    X[:, :n] = Xn; X[:, n] = X[:, kdim]; X[:, n + 1:] = 0
    H[:kdim, :] = T; H[n, :] = b; H[n + 1:, :] = 0; H[:, n:] = 0
    return n


def eigs_literal(apply_A, n, nev, x0, kind, kdim, tol):
    dt = DT[kind]; cplx = kind in "cz"
    Xw = np.zeros((n, kdim + 1), dtype=dt, order="F"); Xw[:, 0] = x0 / v_norm(x0)
    H = np.zeros((kdim + 1, kdim), dtype=dt)
    kstart, conv, niter, k = 1, 0, 0, 0
    while conv < nev:
        for k in range(kstart, kdim + 1):
            Xw[:, k] = apply_A(Xw[:, k - 1])                       # arnoldi(A, Xwrk, H, info, kstart = k, kend = k)
            H[:k, k - 1], _ = dgs_vector(Xw[:, k], Xw[:, :k], kind)
            beta = v_norm(Xw[:, k]); H[k, k - 1] = beta
            Xw[:, k] *= dt(1) / dt(beta)
            if cplx:
                vals, vecs = np.linalg.eig(H[:k, :k])
                res = np.abs(beta * vecs[k - 1, :k])
            else:
                vals, vecs = _eig_real_pair(H[:k, :k])
                res = np.zeros(k)
                for i in range(k):
                    if vals[i].imag > 0:
                        alpha = abs(complex(vecs[k - 1, i], vecs[k - 1, i + 1]))
                    elif vals[i].imag < 0:
                        alpha = abs(complex(vecs[k - 1, i - 1], vecs[k - 1, i]))
                    else:
                        alpha = abs(vecs[k - 1, i])
                    res[i] = abs(beta * alpha)
            niter += 1
            conv = int((res < tol).sum())
            if conv >= nev:
                break
        else:
            k = kdim + 1
        # NOTE the reference restarts ONCE MORE after convergence: `exit arnoldi_factorization` leaves only the inner loop, the
        # krylov_schur call below it still runs before `do while (conv < nev)` is re-evaluated (IterativeSolvers.fypp:1088-1099)
        kstart = _krylov_schur_literal(Xw, H) + 1
    k = min(k, kdim)
    vals = np.linalg.eigvals(H[:k, :k])                            # eig of the RESTARTED matrix (:1108-1110)
    order = np.argsort(-np.abs(vals), kind="stable")           # sort_index(reverse = .true.): non-increasing, ties in original order
    return vals[order[:nev]], niter, k, kstart - 1


@pytest.mark.parametrize("kind", ["d", "z"])
def test_eigs_literal_flow_vs_oracle(oracle, kind):
    """The reference restarts once more AFTER convergence and post-processes the restarted Hessenberg matrix
    (IterativeSolvers.fypp:1088-1117).  Until round 2 the oracle and the product post-processed the converged factorisation
    instead -- this independent restatement of the literal flow exposed the difference; both follow the literal flow now.
    Same info = niter, same eigenvalues (they equal the plain post-processing to rounding whenever k >= n, the size of the
    retained Schur block)."""
    dt = DT[kind]; n, nev, kdim = 120, 4, 24
    rng = np.random.default_rng(9)
    Q, _ = np.linalg.qr(randn(rng, (n, n), dt))
    lam = np.concatenate([[9.0, -8.2, 7.1, 6.3], np.linspace(-3.0, 3.0, n - 4)]).astype(dt)
    Npart = np.triu(randn(rng, (n, n), dt), 1) * 0.02               # non-normal part: a genuinely nonsymmetric operator
    A = np.asfortranarray((Q @ (np.diag(lam) + Npart) @ Q.conj().T).astype(dt))
    x0 = randn(rng, n, dt)
    ev2, niter2, k2, nkeep = eigs_literal(lambda v: (A @ v).astype(dt), n, nev, x0.copy(), kind, kdim, 1e-9)
    evo, reso, _, niter_o = oracle.eigs(oracle.Op.dense(A), n, nev, x0.copy(), kdim=kdim, tolerance=1e-9)
    assert niter2 == niter_o > kdim                                 # several Krylov-Schur cycles, identical iteration count
    assert k2 >= nkeep                                              # the regime the statement above is about
    key = lambda z: (-round(abs(z), 6), round(z.imag, 6))
    a = np.array(sorted(ev2, key=key)); b = np.array(sorted(evo, key=key))
    assert np.abs(a - b).max() < 1e-9 * np.abs(b).max()
    assert np.abs(np.sort(np.abs(b))[::-1] - np.array([9.0, 8.2, 7.1, 6.3])).max() < 1e-6


# ---- kexpm_vec, LITERAL control flow of src/Expm/ExpmLib.fypp:176-230 (dense expm = scipy's here, not the oracle's Pade-10) ----
def kexpm_vec_literal(apply_A, b, tau, tol, kind, nk=100):
    import scipy.linalg as sla
    dt = DT[kind]; n = b.size
    beta = v_norm(b)
    X = np.zeros((n, nk + 1), dtype=dt, order="F"); X[:, 0] = b / dt(beta)
    H = np.zeros((nk + 1, nk + 1), dtype=dt)
    c = np.zeros(n, dtype=dt); err_est = 0.0; kp = 1
    rng = np.random.default_rng(99)
    for k in range(1, nk + 1):
        kp = k + 1
        # arnoldi(A, X, H, info, kstart = k, kend = k): one step, incl. qr_no_pivoting's refill on breakdown (qr.fypp:146-162)
        X[:, k] = apply_A(X[:, k - 1])
        H[:k, k - 1], _ = dgs_vector(X[:, k], X[:, :k], kind)
        bk = v_norm(X[:, k]); info = 0
        if bk < ATOL[kind]:
            H[k, k - 1] = 0
            X[:, k] = randn(rng, n, dt)
            dgs_vector(X[:, k], X[:, :k], kind)
            X[:, k] /= dt(v_norm(X[:, k]))
            info = k                                              # arnoldi: |H(k+1,k)| < tol => info = k
        else:
            H[k, k - 1] = bk
            X[:, k] *= dt(1) / dt(bk)
        if info == k:
            kp = k
            info = -2
        E = sla.expm(tau * H[:kp, :kp])
        c = (dt(beta) * (X[:, :kp] @ E[:kp, 0])).astype(dt)
        err_est = 0.0 if info == k else abs(E[kp - 1, 0] * beta)   # merge(0, abs(E(kp,1)*beta), info == k): info is -2 by now
        if err_est <= tol:
            break
    return c, (kp if err_est <= tol else -1)


@pytest.mark.parametrize("kind", ["d", "z"])
def test_kexpm_vec_two_restatements_agree(oracle, kind):
    import scipy.linalg as sla
    dt = DT[kind]; n = 90
    rng = np.random.default_rng(10)
    A = np.asfortranarray((randn(rng, (n, n), dt) / np.sqrt(n) - 0.5 * np.eye(n)).astype(dt))
    b = randn(rng, n, dt)
    c2, info2 = kexpm_vec_literal(lambda v: (A @ v).astype(dt), b, 0.3, 1e-10, kind)
    co, infoo = oracle.kexpm_vec(oracle.Op.dense(A), b, 0.3, 1e-10)
    assert info2 == infoo and 1 < info2 < 60
    assert np.linalg.norm(c2 - co) < 1e-10 * np.linalg.norm(co)
    assert np.linalg.norm(co - sla.expm(0.3 * A) @ b) < 1e-8 * np.linalg.norm(co)
    # breakdown: b inside a 3-dimensional invariant subspace of a diagonal operator.  Literal behaviour: the estimate is not
    # zeroed at the breakdown step k = 3 (|E(3,1) beta| ~ 0.1 > tol), the loop continues on the refilled vector and stops at
    # step 4 with kp = 5, where E(kp,1) is exactly zero -- both restatements, same info, same (exact) vector
    D = np.asfortranarray(np.diag(np.arange(1, n + 1)).astype(dt))
    b3 = np.zeros(n, dtype=dt); b3[:3] = [1.0, 2.0, -1.5]
    c2, info2 = kexpm_vec_literal(lambda v: (D @ v).astype(dt), b3, 0.2, 1e-12, kind)
    co, infoo = oracle.kexpm_vec(oracle.Op.dense(D), b3, 0.2, 1e-12)
    assert info2 == infoo == 5
    exact = np.exp(0.2 * np.arange(1, n + 1)) * b3
    assert np.linalg.norm(c2 - exact) < 1e-12 * np.linalg.norm(exact) and np.linalg.norm(co - exact) < 1e-12 * np.linalg.norm(exact)


@pytest.mark.parametrize("mode", ["right-preconditioned gmres", "fgmres"])
def test_preconditioned_gmres_two_restatements_agree(oracle, mode):
    """gmres with a (fixed) right preconditioner and fgmres with a preconditioner that changes with the inner step k
    (gmres.fypp:153-156, 205-206; fgmres.fypp:158-161, 205-206)."""
    kind, dt, n = "d", np.float64, 180
    rng = np.random.default_rng(11)
    dg = np.linspace(1.0, 40.0, n)
    A = np.asfortranarray(np.diag(dg) + 0.4 * randn(rng, (n, n), dt))
    b = randn(rng, n, dt)
    flexible = mode == "fgmres"
    scale = (lambda k: 1.0 + 0.2 * (k % 3)) if flexible else (lambda k: 1.0)
    prec2 = lambda v, k: (v / dg) * scale(k)

    def prec_o(v, k=0):
        v[:] = (v / dg) * scale(k)
    x2 = np.zeros(n); xo = np.zeros(n)
    info2, m2 = gmres(lambda v: A @ v, b, x2, kind, kdim=10, maxiter=40, precond=prec2, flexible=flexible)
    infoo, mo = oracle.gmres(oracle.Op.dense(A), b, xo, kdim=10, maxiter=40, precond=prec_o, flexible=flexible)
    assert info2 == infoo > 0 and m2["n_outer"] == mo["n_outer"] >= 2
    assert np.abs(np.array(m2["res"]) - np.array(mo["res"])).max() < 1e-10 * mo["res"][0]
    assert np.linalg.norm(x2 - xo) < 1e-10 * np.linalg.norm(xo)
    assert np.linalg.norm(A @ x2 - b) < 1e-6 * np.linalg.norm(b)


# ---- qr_with_pivoting (src/Krylov/qr.fypp:32-107, swap_columns :174-201) and kexpm_mat (src/Expm/ExpmLib.fypp:234-362) -------
def qr_with_pivoting_literal(Q, kind, tol=None):
    """Written from the Fortran with 1-based loop variables, no breakdown branch exercised except the rank-exhausted exit."""
    dt = Q.dtype
    tolerance = ATOL[kind] if tol is None else tol
    kdim = Q.shape[1]
    R = np.zeros((kdim, kdim), dtype=dt)
    perm = np.zeros(kdim, dtype=int)
    Rii = np.zeros(kdim, dtype=dt)
    info = 0
    rng = np.random.default_rng(123)
    for i in range(1, kdim + 1):
        perm[i - 1] = i
        Rii[i - 1] = v_dot(Q[:, i - 1], Q[:, i - 1])
    for j in range(1, kdim + 1):
        idx = int(np.argmax(np.abs(Rii))) + 1                       # maxloc(abs(Rii))
        if abs(Rii[idx - 1]) < tolerance:
            for i in range(j, kdim + 1):
                Q[:, i - 1] = randn(rng, Q.shape[0], dt)
                if i > 1:
                    dgs_vector(Q[:, i - 1], Q[:, :i - 1], kind)
                Q[:, i - 1] *= dt.type(1) / dt.type(v_norm(Q[:, i - 1]))
            info = j
            break
        # swap_columns(Q, R, Rii, perm, j, idx): n = min(j, idx) - 1 = j - 1 leading rows of R
        Q[:, [j - 1, idx - 1]] = Q[:, [idx - 1, j - 1]]
        Rii[[j - 1, idx - 1]] = Rii[[idx - 1, j - 1]]
        perm[[j - 1, idx - 1]] = perm[[idx - 1, j - 1]]
        n = min(j, idx) - 1
        if n > 0:
            R[:n, [j - 1, idx - 1]] = R[:n, [idx - 1, j - 1]]
        beta = v_norm(Q[:, j - 1])
        assert np.isfinite(beta) and beta >= tolerance, "these inputs must not cancel"
        R[j - 1, j - 1] = beta
        Q[:, j - 1] *= dt.type(1) / dt.type(beta)
        for i in range(j + 1, kdim + 1):
            b = v_dot(Q[:, j - 1], Q[:, i - 1])
            v_axpby(-b, Q[:, j - 1], 1, Q[:, i - 1])
            R[j - 1, i - 1] = b
        Rii[j - 1] = 0
        for i in range(j + 1, kdim + 1):
            Rii[i - 1] = Rii[i - 1] - R[j - 1, i - 1] ** 2
    return R, perm, info


@pytest.mark.parametrize("kind", ["s", "d", "c", "z"])
def test_qr_with_pivoting_two_restatements_agree(oracle, kind):
    dt = DT[kind]
    rng = np.random.default_rng(40)
    A = randn(rng, (100, 14), dt)
    A[:, 5] = 0; A[:, 11] = 0                                        # exact rank deficiency 2: exit at step 13
    Q2 = A.copy(order="F"); R2, perm2, info2 = qr_with_pivoting_literal(Q2, kind)
    Qo = A.copy(order="F"); infoo, Ro, permo = oracle.qr_with_pivoting(Qo)
    assert info2 == infoo == 13
    assert (perm2 - 1).tolist() == permo.tolist()                    # the oracle returns perm 0-based
    assert rel_normwise(R2, Ro) < _tol(kind) and rel_normwise(Q2[:, :12], Qo[:, :12]) < _tol(kind)
    for Q, R, perm in ((Q2, R2, perm2 - 1), (Qo, Ro, permo)):
        assert np.abs(A[:, perm] - Q @ R).max() < 100 * _tol(kind)
        assert np.abs(Q.conj().T @ Q - np.eye(14)).max() < 100 * _tol(kind)


def kexpm_mat_literal(apply_A, B, tau, tol, kind, kdim=100):
    """ExpmLib.fypp:234-362 with its own pivoting QR, block Arnoldi step (arnoldi.fypp:34-73 for one k) and scipy's expm."""
    import scipy.linalg as sla
    dt = DT[kind]
    n, p = B.shape
    nsteps = kdim; nk = nsteps * p
    Xwrk = B.copy(order="F")
    R, perm, _ = qr_with_pivoting_literal(Xwrk, kind)
    inv_perm = np.zeros(p, dtype=int); inv_perm[perm - 1] = np.arange(1, p + 1)         # invperm (utilities.fypp:21-27)
    R = R[:, inv_perm - 1]                                                              # permcols(R, invperm(perm))
    if np.sqrt((np.abs(R) ** 2).sum()) == 0:                                            # mnorm(R, "fro") == 0
        return np.zeros((n, p), dtype=dt), p
    X = np.zeros((n, p * (nk + 1)), dtype=dt, order="F")
    X[:, :p] = Xwrk
    qr_no_pivoting(X[:, :p], kind)                                                      # initialize_krylov_subspace
    H = np.zeros((p * (nk + 1), p * (nk + 1)), dtype=dt)
    C = np.zeros((n, p), dtype=dt); err_est = 0.0; kpp = p
    for k in range(1, nk + 1):
        kpm = (k - 1) * p; kp = kpm + p; kpp = kp + p
        # arnoldi(A, X, H, info, kstart = k, kend = k, blksize = p)
        for i in range(p):
            X[:, kp + i] = apply_A(X[:, kpm + i])
        if p == 1:
            H[:kp, kpm], _ = dgs_vector(X[:, kp], X[:, :kp], kind)
        else:
            H[:kp, kpm:kp], _ = dgs_basis(X[:, kp:kpp], X[:, :kp], kind)
        Rb, _ = qr_no_pivoting(X[:, kp:kpp], kind)                                      # (no breakdown in these inputs)
        H[kp:kpp, kpm:kp] = Rb
        E = sla.expm(tau * H[:kpp, :kpp])
        Xw = np.zeros((n, p), dtype=dt)
        for i in range(p):
            for j in range(kpp):
                v_axpby(E[j, i], X[:, j], 1, Xw[:, i])
        C[:] = 0
        for i in range(p):
            for j in range(p):
                v_axpby(R[j, i], Xw[:, j], 1, C[:, i])
        err_est = np.sqrt((np.abs(E[kp:kpp, :p] @ R[:p, :p]) ** 2).sum())               # norm(matmul(...), 2): all elements
        if err_est <= tol:
            break
    return C, (kpp if err_est <= tol else -1)


@pytest.mark.parametrize("kind", ["d", "z"])
@pytest.mark.parametrize("p", [1, 2, 3])
def test_kexpm_mat_two_restatements_agree(oracle, kind, p):
    import scipy.linalg as sla
    dt = DT[kind]; n = 90
    rng = np.random.default_rng(41)
    A = np.asfortranarray((randn(rng, (n, n), dt) / np.sqrt(n) - 0.5 * np.eye(n)).astype(dt))
    B = randn(rng, (n, p), dt)
    C2, info2 = kexpm_mat_literal(lambda v: (A @ v).astype(dt), B, 0.3, 1e-10, kind, kdim=12)
    Co, infoo = oracle.kexpm_mat(oracle.Op.dense(A), B, 0.3, 1e-10, kdim=12)
    assert info2 == infoo and info2 > p
    assert np.linalg.norm(C2 - Co) < 1e-10 * np.linalg.norm(Co)
    assert np.linalg.norm(Co - sla.expm(0.3 * A) @ B) < 1e-8 * np.linalg.norm(Co)
