"""ctypes front-end of the CPU oracle (oracle/lk_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product package (lightkrylov_b200) never does.

The C library restates the reference's per-vector algorithm (file:line map in
lko_body.inc).  This module adds thin numpy marshalling plus the *solver shells*
(gmres / cg / eigs / eighs / svds / krylov_schur) restated in numpy + scipy LAPACK
on top of the C primitives, following
  src/IterativeSolvers/GMRES/gmres.fypp:65-255
  src/IterativeSolvers/CG/CG.fypp:61-196
  src/IterativeSolvers/IterativeSolvers.fypp:972-1143   (eigs)
  src/Krylov/BaseKrylov.fypp:782-834                    (krylov_schur)
  src/IterativeSolvers/EIGHS/eighs.fypp:29-126
  src/IterativeSolvers/SVDS/svd_solvers.fypp:28-121
  src/Utilities/submodule_utility_functions.fypp:55-117, 169-204
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

DTYPES = {"s": np.float32, "d": np.float64, "c": np.complex64, "z": np.complex128}
REAL = {"s": np.float32, "d": np.float64, "c": np.float32, "z": np.float64}
ATOL = {"s": 1e-6, "d": 1e-15, "c": 1e-6, "z": 1e-15}          # Constants.f90:16-37
RTOL = {k: float(np.sqrt(v)) for k, v in ATOL.items()}


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liblk_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("lk_oracle.c", "lko_body.inc", "Makefile")]
    stale = (not os.path.exists(so)) or any(
        os.path.exists(s) and os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "liblk_oracle.so"],
                              stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.lko_max_threads.restype = C.c_int
    return _LIB


def set_threads(n: int) -> None:
    lib().lko_set_threads(C.c_int(int(n)))


def max_threads() -> int:
    return int(lib().lko_max_threads())


def kind_of(dtype) -> str:
    dt = np.dtype(dtype)
    for k, v in DTYPES.items():
        if np.dtype(v) == dt:
            return k
    raise TypeError(f"unsupported dtype {dt}")


class _OpStruct(C.Structure):
    """Mirror of lko_op_X; coef is 7 elements of T so the layout depends on the kind."""


def _op_struct(kind: str):
    esz = np.dtype(DTYPES[kind]).itemsize

    class S(C.Structure):
        _fields_ = [("kind", C.c_int), ("m", C.c_int64), ("n", C.c_int64), ("a", C.c_void_p),
                    ("nx", C.c_int64), ("ny", C.c_int64), ("nz", C.c_int64),
                    ("coef", C.c_ubyte * (7 * esz)),
                    ("rowptr", C.c_void_p), ("col", C.c_void_p),
                    ("n_matvec", C.c_int64), ("n_rmatvec", C.c_int64)]
    return S


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class Op:
    """Operator handle for the oracle: dense / stencil5 / stencil7 / csr."""

    def __init__(self, kind: str, st, keep, m, n):
        self.kind, self.st, self._keep, self.m, self.n = kind, st, keep, m, n

    @staticmethod
    def dense(A: np.ndarray) -> "Op":
        kind = kind_of(A.dtype)
        Af = np.asfortranarray(A)
        st = _op_struct(kind)()
        st.kind, st.m, st.n, st.a = 0, A.shape[0], A.shape[1], Af.ctypes.data
        return Op(kind, st, [Af], A.shape[0], A.shape[1])

    @staticmethod
    def stencil(kind: str, dims, coef) -> "Op":
        """coef = (center, -x, +x, -y, +y[, -z, +z]); grid index i + nx*(j + ny*k)."""
        st = _op_struct(kind)()
        dims = list(dims)
        c = np.zeros(7, dtype=DTYPES[kind])
        c[:len(coef)] = coef
        C.memmove(st.coef, c.ctypes.data, c.nbytes)
        if len(dims) == 2:
            st.kind, st.nx, st.ny, st.nz = 1, dims[0], dims[1], 1
        else:
            st.kind, st.nx, st.ny, st.nz = 2, dims[0], dims[1], dims[2]
        n = int(np.prod(dims))
        st.m = st.n = n
        return Op(kind, st, [c], n, n)

    @staticmethod
    def csr(m, n, rowptr, col, val) -> "Op":
        kind = kind_of(val.dtype)
        rp = np.ascontiguousarray(rowptr, dtype=np.int64)
        ci = np.ascontiguousarray(col, dtype=np.int32)
        va = np.ascontiguousarray(val)
        st = _op_struct(kind)()
        st.kind, st.m, st.n, st.a = 3, m, n, va.ctypes.data
        st.rowptr, st.col = rp.ctypes.data, ci.ctypes.data
        return Op(kind, st, [rp, ci, va], m, n)

    def apply(self, x: np.ndarray, trans: bool = False) -> np.ndarray:
        y = np.empty(self.n if trans else self.m, dtype=DTYPES[self.kind])
        x = np.ascontiguousarray(x, dtype=DTYPES[self.kind])
        getattr(lib(), f"lko_apply_{self.kind}")(C.byref(self.st), _ptr(x), _ptr(y), C.c_int(int(trans)))
        return y

    @property
    def n_matvec(self):
        return int(self.st.n_matvec)


def _real_arg(kind, v):
    return C.c_float(v) if REAL[kind] == np.float32 else C.c_double(v)


def fill(n: int, kind: str, dist: str, seed: int, row0: int = 0) -> np.ndarray:
    """Counter-based random vector, bit-identical to the device generator for 'uniform'."""
    x = np.empty(n, dtype=DTYPES[kind])
    getattr(lib(), f"lko_fill_{kind}")(C.c_int64(n), _ptr(x), C.c_int(0 if dist == "normal" else 1),
                                       C.c_uint64(seed), C.c_int64(row0))
    return x


def normalize(x: np.ndarray) -> None:
    getattr(lib(), f"lko_normalize_{kind_of(x.dtype)}")(C.c_int64(x.size), _ptr(x))


def dot(x, y):
    kind = kind_of(x.dtype)
    out = np.zeros(1, dtype=DTYPES[kind])
    getattr(lib(), f"lko_dot_{kind}")(C.c_int64(x.size), _ptr(x), _ptr(y), _ptr(out))
    return out[0]


def axpby(alpha, x, beta, y):
    """y = alpha*x + beta*y in place."""
    kind = kind_of(y.dtype)
    a = np.array([alpha], dtype=DTYPES[kind]); b = np.array([beta], dtype=DTYPES[kind])
    getattr(lib(), f"lko_axpby_{kind}")(C.c_int64(y.size), _ptr(a), _ptr(x), _ptr(b), _ptr(y))


def norm(x):
    return float(np.sqrt(np.abs(dot(x, x))))


def _check_basis(X: np.ndarray):
    assert X.flags.f_contiguous and X.ndim == 2, "basis must be Fortran-ordered 2-D"


def dgs_vec(y: np.ndarray, X: np.ndarray, j: int, want_beta: bool = True):
    """double_gram_schmidt_step(y, X(:j)) -> (info, beta)."""
    _check_basis(X)
    kind = kind_of(X.dtype)
    beta = np.zeros(max(j, 1), dtype=X.dtype)
    info = getattr(lib(), f"lko_dgs_vec_{kind}")(C.c_int64(X.shape[0]), _ptr(y), _ptr(X), C.c_int64(X.shape[0]),
                                                C.c_int(j), _ptr(beta) if want_beta else None)
    return int(info), beta[:j]


def dgs_bas(Y: np.ndarray, X: np.ndarray, j: int):
    _check_basis(X); _check_basis(Y)
    kind = kind_of(X.dtype)
    p = Y.shape[1]
    beta = np.zeros((max(j, 1), p), dtype=X.dtype, order="F")
    info = getattr(lib(), f"lko_dgs_bas_{kind}")(C.c_int64(X.shape[0]), _ptr(Y), C.c_int64(Y.shape[0]), C.c_int(p),
                                                _ptr(X), C.c_int64(X.shape[0]), C.c_int(j), _ptr(beta),
                                                C.c_int(beta.shape[0]))
    return int(info), beta[:j]


def qr(Q: np.ndarray, tol=None, seed: int = 1000):
    _check_basis(Q)
    kind = kind_of(Q.dtype)
    p = Q.shape[1]
    Rm = np.zeros((p, p), dtype=Q.dtype, order="F")
    sd = C.c_uint64(seed)
    info = getattr(lib(), f"lko_qr_{kind}")(C.c_int64(Q.shape[0]), _ptr(Q), C.c_int64(Q.shape[0]), C.c_int(p),
                                           _ptr(Rm), C.c_int(p), _real_arg(kind, ATOL[kind] if tol is None else tol),
                                           C.byref(sd), C.c_int64(0))
    return int(info), Rm


def arnoldi(A: Op, X: np.ndarray, H: np.ndarray, kstart=1, kend=None, tol=None, trans=False,
            blksize=1, seed=1000):
    """arnoldi(A, X, H, info, kstart, kend, tol, transpose, blksize); X, H updated in place."""
    _check_basis(X); _check_basis(H)
    kind = kind_of(X.dtype)
    kdim = (X.shape[1] - blksize) // blksize
    kend = kdim if kend is None else kend
    sd = C.c_uint64(seed)
    info = getattr(lib(), f"lko_arnoldi_{kind}")(
        C.byref(A.st), C.c_int64(X.shape[0]), _ptr(X), C.c_int64(X.shape[0]), _ptr(H), C.c_int(H.shape[0]),
        C.c_int(kdim), C.c_int(kstart), C.c_int(kend), _real_arg(kind, ATOL[kind] if tol is None else tol),
        C.c_int(int(trans)), C.c_int(blksize), C.byref(sd))
    return int(info)


def lanczos(A: Op, X: np.ndarray, T: np.ndarray, kstart=1, kend=None, tol=None):
    _check_basis(X); _check_basis(T)
    kind = kind_of(X.dtype)
    kdim = X.shape[1] - 1
    kend = kdim if kend is None else kend
    info = getattr(lib(), f"lko_lanczos_{kind}")(
        C.byref(A.st), C.c_int64(X.shape[0]), _ptr(X), C.c_int64(X.shape[0]), _ptr(T), C.c_int(T.shape[0]),
        C.c_int(kdim), C.c_int(kstart), C.c_int(kend), _real_arg(kind, ATOL[kind] if tol is None else tol))
    return int(info)


def bidiag(A: Op, U: np.ndarray, V: np.ndarray, B: np.ndarray, kstart=1, kend=None, tol=None):
    _check_basis(U); _check_basis(V); _check_basis(B)
    kind = kind_of(U.dtype)
    kdim = U.shape[1] - 1
    kend = kdim if kend is None else kend
    info = getattr(lib(), f"lko_bidiag_{kind}")(
        C.byref(A.st), C.c_int64(U.shape[0]), C.c_int64(V.shape[0]), _ptr(U), C.c_int64(U.shape[0]),
        _ptr(V), C.c_int64(V.shape[0]), _ptr(B), C.c_int(B.shape[0]), C.c_int(kdim),
        C.c_int(kstart), C.c_int(kend), _real_arg(kind, ATOL[kind] if tol is None else tol))
    return int(info)


# ------------------------------------------------------------------------------------------
# Solver shells (numpy + scipy LAPACK over the C primitives)
# ------------------------------------------------------------------------------------------

def _lartg(f, g):
    """LAPACK 3.10 la_lartg (real): c = |f|/d, r = sign(d, f), s = g/r."""
    if g == 0:
        return 1.0, 0.0, f
    if f == 0:
        return 0.0, float(np.sign(g)), abs(g)
    d = np.hypot(f, g)
    c = abs(f) / d
    r = np.copysign(d, f)
    return c, g / r, r


def apply_givens_rotation(h, c, s):
    """submodule_utility_functions.fypp:174-204; h has k+1 entries, c/s have k entries."""
    k = h.size - 1
    if np.iscomplexobj(h):
        for i in range(k - 1):
            t = c[i] * h[i] + s[i] * h[i + 1]
            h[i + 1] = -s[i] * h[i] + c[i] * h[i + 1]
            h[i] = t
        nrm = np.sqrt(abs(h[k - 1]) ** 2 + abs(h[k]) ** 2)     # g = x / norm(x, 2)
        c[k - 1], s[k - 1] = h[k - 1] / nrm, h[k] / nrm
        h[k - 1] = c[k - 1] * h[k - 1] + s[k - 1] * h[k]
        h[k] = 0
    else:
        for j in range(k - 1):                                  # lasr('L','V','F')
            t = h[j + 1]
            h[j + 1] = c[j] * t - s[j] * h[j]
            h[j] = s[j] * t + c[j] * h[j]
        c[k - 1], s[k - 1], r = _lartg(h[k - 1], h[k])
        h[k - 1] = r
        h[k] = 0


def gmres(A: Op, b: np.ndarray, x: np.ndarray, rtol=None, atol=None, kdim=30, maxiter=10, trans=False):
    """gmres.fypp:65-255 (no preconditioner).  Returns (info, meta dict); x updated in place."""
    import scipy.linalg as sla
    kind = kind_of(b.dtype)
    n = b.size
    rtol = RTOL[kind] if rtol is None else rtol
    atol = ATOL[kind] if atol is None else atol
    tol = atol + rtol * norm(b)
    V = np.zeros((n, kdim + 1), dtype=b.dtype, order="F")
    meta = dict(n_iter=0, n_inner=0, n_outer=0, res=[], converged=False)
    while (not meta["converged"]) and meta["n_outer"] <= maxiter:
        H = np.zeros((kdim + 1, kdim), dtype=b.dtype, order="F")
        V[:] = 0
        if norm(x) != 0.0:
            V[:, 0] = A.apply(x, trans)
        axpby(-1, b, 1, V[:, 0]); V[:, 0] *= -1
        e = np.zeros(kdim + 1, dtype=b.dtype)
        beta = norm(V[:, 0]); e[0] = beta
        V[:, 0] *= (1.0 / beta)
        c = np.zeros(kdim, dtype=b.dtype); s = np.zeros(kdim, dtype=b.dtype)
        if meta["n_outer"] == 0:
            meta["res"].append(abs(beta))
        kk = kdim
        for k in range(1, kdim + 1):
            V[:, k] = A.apply(V[:, k - 1].copy(), trans)
            _, hcol = dgs_vec(V[:, k], V, k)
            H[:k, k - 1] = hcol
            H[k, k - 1] = norm(V[:, k])
            if abs(H[k, k - 1]) > tol:
                V[:, k] *= (1.0 / H[k, k - 1])
            hv = H[:k + 1, k - 1].copy()
            apply_givens_rotation(hv, c[:k], s[:k])
            H[:k + 1, k - 1] = hv
            e[k] = -s[k - 1] * e[k - 1]; e[k - 1] = c[k - 1] * e[k - 1]
            beta = abs(e[k])
            meta["n_iter"] += 1; meta["n_inner"] += 1; meta["res"].append(abs(beta))
            if abs(beta) < tol:
                meta["converged"] = True
                kk = k
                break
        k = kk
        y = sla.solve_triangular(H[:k, :k], e[:k], lower=False)
        dx = V[:, :k] @ y
        x += dx
        V[:, 0] = A.apply(x, trans)
        axpby(-1, b, 1, V[:, 0]); V[:, 0] *= -1
        beta = norm(V[:, 0])
        meta["n_iter"] += 1; meta["n_outer"] += 1; meta["res"].append(abs(beta))
        if abs(beta) < tol:
            meta["converged"] = True
            break
    info = meta["n_iter"] if meta["converged"] else -meta["n_iter"]
    meta["info"] = info
    return info, meta


def cg(A: Op, b: np.ndarray, x: np.ndarray, rtol=None, atol=None, maxiter=100):
    """CG.fypp:61-196 (no preconditioner)."""
    kind = kind_of(b.dtype)
    rtol = RTOL[kind] if rtol is None else rtol
    atol = ATOL[kind] if atol is None else atol
    tol = atol + rtol * norm(b)
    r = np.zeros_like(b)
    if norm(x) > 0:
        r = A.apply(x)
    axpby(-1, b, 1, r); r *= -1
    p = r.copy()
    rr_old = dot(r, r)
    meta = dict(n_iter=0, res=[float(np.sqrt(abs(rr_old)))], converged=False)
    for _ in range(maxiter):
        Ap = A.apply(p)
        alpha = rr_old / dot(p, Ap)
        axpby(alpha, p, 1, x)
        axpby(-alpha, Ap, 1, r)
        rr_new = dot(r, r)
        residual = float(np.sqrt(abs(rr_new)))
        meta["n_iter"] += 1; meta["res"].append(residual)
        if residual < tol:
            meta["converged"] = True
            break
        beta = rr_new / rr_old
        axpby(1, r, beta, p)
        rr_old = rr_new
    info = meta["n_iter"] if meta["converged"] else -meta["n_iter"]
    return info, meta
