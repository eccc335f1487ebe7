"""A SECOND, independent restatement of the reference's step loops, used only to cross-check the C oracle (oracle/).

The C oracle (oracle/lk_oracle.c + lko_body.inc) is what every GPU parity test compares against, and nothing in this
environment can execute the Fortran reference (no compiler here or on the GPU box: parity is unpinned against an execution
of the reference).  To reduce the risk that the oracle and the product share one misreading of the
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
