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


def sym_matrix(kind, seed=3):
    M = pseudo((N, N), seed, kind).astype(np.complex128 if kind in "cz" else np.float64)
    S = M @ M.conj().T / N + 0.01 * np.eye(N)
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


CASES = {
    "arnoldi_full": arnoldi_full, "arnoldi_transpose": arnoldi_transpose, "arnoldi_block": arnoldi_block,
    "arnoldi_resume": arnoldi_resume, "arnoldi_breakdown": arnoldi_breakdown, "lanczos_full": lanczos_full,
    "bidiag_full": bidiag_full, "qr_full": qr_full, "qr_deficient": qr_deficient, "qr_pivoting": qr_pivoting,
    "dgs_vector": dgs_vector, "dgs_basis": dgs_basis,
}


SOLVER_CASES = {}


def applies(name, kind):
    return True


# ------------------------------------------------------------------------------------------------------------------ backends
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

    def lanczos(self, A, X, T):
        _, o = self.rx.call("lanczos", A, X, T, 0)
        return int(o[3])

    def bidiag(self, A, U, V, B):
        _, o = self.rx.call("bidiagonalization", A, U, V, B, 0)
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

    def counter(self, A):
        return int(A.n_matvec)

    def arnoldi(self, A, X, H, kstart=None, kend=None, tol=None, transpose=None, blksize=None):
        return self.lo.arnoldi(A, X, H, kstart=kstart or 1, kend=kend, tol=tol, trans=bool(transpose), blksize=blksize or 1)

    def lanczos(self, A, X, T):
        return self.lo.lanczos(A, X, T)

    def bidiag(self, A, U, V, B):
        return self.lo.bidiag(A, U, V, B)

    def qr(self, Q, tol=None):
        return self.lo.qr(Q, tol=tol)

    def qr_pivoting(self, Q):
        return self.lo.qr_with_pivoting(Q)

    def dgs_vec(self, y, X):
        return self.lo.dgs_vec(y[:, 0], X, X.shape[1])

    def dgs_bas(self, Y, X):
        return self.lo.dgs_bas(Y, X, X.shape[1])
