"""lkb_mock -- liblkb's C ABI (include/lkb.h) on HOST memory, as natives of oracle/f90run.py.  TEST INFRASTRUCTURE ONLY.

Purpose: the Fortran shim fortran/lightkrylov_cuda.f90 can be compiled nowhere (no Fortran compiler in this image or on the GPU
box).  With this mock the interpreter EXECUTES it: its type-bound procedures (zero / rand / scal / axpby / dot / get_size,
defined assignment, lazy allocation, the magic-number liveness test), its constructors (cuda_basis_allocate_*, cuda_csr_*) and
its `lkb_try_*` dispatchers run against "device" objects that are numpy arrays, the Krylov entry points being served by the C
oracle (oracle/lk_oracle.py).  tests/test_shim_under_interpreter.py then drives the REFERENCE's own arnoldi / lanczos / qr /
gmres ... (also interpreted) with the shim's types, both through the dispatchers and through the reference's generic loop.

What this checks: the Fortran side of the boundary (argument order and by-value-ness at every call site, handle lifetime, view
bookkeeping, info / optional-argument translation).  What it cannot check: the CUDA library itself (that is `-m gpu`).
Only the entry points the shim declares are provided; every handle is an opaque Python object inside a CPtr.
"""
import numpy as np

from . import lk_oracle as lo
from .f90run import CPtr, FortranError, ScalarRef

KINDS = "sdcz"                      # LKB_S = 0, LKB_D = 1, LKB_C = 2, LKB_Z = 3
DT = {"s": np.float32, "d": np.float64, "c": np.complex64, "z": np.complex128}
LKB_OK, LKB_ERR_ARG = 0, -1


class MVec:
    def __init__(self, kind, data, owner=None):
        self.kind, self.data, self.owner, self.alive = kind, data, owner, True


class MBasis:
    def __init__(self, kind, data, parent=None):
        self.kind, self.data, self.parent, self.alive = kind, data, parent, True


class MOp:
    def __init__(self, kind, A=None, op=None):
        self.kind = kind
        self.op = op if op is not None else lo.Op.dense(np.asfortranarray(A))
        self.n_matvec = self.n_rmatvec = 0


class Stats:
    calls = {}
    options = {}


def _deref(p):
    """value behind a const void* scalar argument"""
    o = p.obj
    return o.get() if isinstance(o, ScalarRef) else o


def _store(p, v):
    o = p.obj
    if isinstance(o, ScalarRef):
        o.set(v)
    elif isinstance(o, np.ndarray):
        o.ravel(order="F")[0] = v
    else:
        raise FortranError("mock: output pointer is not writable")


def _count(name):
    Stats.calls[name] = Stats.calls.get(name, 0) + 1


def _hostmat(p, ld, ncols, kind):
    """Fortran-ordered (ld x ncols) view of the host matrix behind a void*"""
    a = p.obj
    if not isinstance(a, np.ndarray) or a.dtype != DT[kind]:
        raise FortranError(f"mock: host matrix of kind {kind} expected, got {getattr(a, 'dtype', type(a))}")
    flat = a.reshape(-1, order="F")
    return flat[:ld * ncols].reshape((ld, ncols), order="F")


def _live(h, cls):
    o = h.obj
    if not isinstance(o, cls) or not o.alive:
        raise FortranError(f"mock: dead or foreign handle where a {cls.__name__} is expected")
    return o


# ---------------------------------------------------------------------------------------------- context / options
def lkb_init(interp, device, ctx):
    return ("__out__", {1: CPtr(object())}, LKB_OK)


def lkb_finalize(interp, ctx):
    return LKB_OK


def lkb_set_option(interp, ctx, name, value):
    _count("lkb_set_option")
    chars = name.obj
    text = "".join(str(c) for c in chars.tolist()) if isinstance(chars, np.ndarray) else str(chars)
    Stats.options[text.split("\x00")[0]] = int(value)
    return LKB_OK


# ---------------------------------------------------------------------------------------------- vectors
def lkb_vec_create(interp, ctx, kind, n_local, n_global, row0, v):
    _count("lkb_vec_create")
    k = KINDS[int(kind)]
    return ("__out__", {5: CPtr(MVec(k, np.zeros(int(n_local), dtype=DT[k])))}, LKB_OK)


def lkb_vec_clone(interp, src, dst):
    s = _live(src, MVec)
    return ("__out__", {1: CPtr(MVec(s.kind, s.data.copy()))}, LKB_OK)


def lkb_vec_destroy(interp, v):
    _count("lkb_vec_destroy")
    _live(v, MVec).alive = False
    return LKB_OK


def lkb_vec_zero(interp, v):
    _live(v, MVec).data[...] = 0
    return LKB_OK


def lkb_vec_rand(interp, v, ifnorm):
    m = _live(v, MVec)
    r = interp.rng.standard_normal(m.data.shape)
    if m.kind in "cz":
        r = r + 1j * interp.rng.standard_normal(m.data.shape)
    m.data[...] = r
    if int(ifnorm):
        m.data[...] /= np.linalg.norm(m.data)
    return LKB_OK


def lkb_vec_scal(interp, v, alpha):
    m = _live(v, MVec)
    m.data[...] = m.data * DT[m.kind](_deref(alpha))
    return LKB_OK


def lkb_vec_axpby(interp, alpha, x, beta, self_):
    xs, ys = _live(x, MVec), _live(self_, MVec)
    if xs.kind != ys.kind or xs.data.shape != ys.data.shape:
        return LKB_ERR_ARG
    a, b = DT[ys.kind](_deref(alpha)), DT[ys.kind](_deref(beta))
    ys.data[...] = a * xs.data if b == 0 else a * xs.data + b * ys.data          # beta == 0 overwrites without reading self
    return LKB_OK


def lkb_vec_dot(interp, self_, vec, out):
    a, b = _live(self_, MVec), _live(vec, MVec)
    _store(out, lo.dot(np.ascontiguousarray(a.data), np.ascontiguousarray(b.data)))
    return LKB_OK


def lkb_vec_size(interp, v):
    return int(_live(v, MVec).data.shape[0])


def lkb_vec_put(interp, v, host):
    m = _live(v, MVec)
    m.data[...] = host.obj.reshape(-1, order="F")[:m.data.shape[0]]
    return LKB_OK


def lkb_vec_get(interp, v, host):
    m = _live(v, MVec)
    host.obj.reshape(-1, order="F")[:m.data.shape[0]] = m.data
    return LKB_OK


# ---------------------------------------------------------------------------------------------- bases
def lkb_basis_create(interp, ctx, kind, n_local, n_global, row0, ncols, b):
    _count("lkb_basis_create")
    k = KINDS[int(kind)]
    return ("__out__", {6: CPtr(MBasis(k, np.zeros((int(n_local), int(ncols)), dtype=DT[k], order="F")))}, LKB_OK)


def lkb_basis_destroy(interp, b):
    _count("lkb_basis_destroy")
    _live(b, MBasis).alive = False
    return LKB_OK


def lkb_basis_col(interp, b, i0, view):
    m = _live(b, MBasis)
    if not 0 <= int(i0) < m.data.shape[1]:
        return LKB_ERR_ARG
    return ("__out__", {2: CPtr(MVec(m.kind, m.data[:, int(i0)], owner=m))}, LKB_OK)


def lkb_basis_view(interp, b, col0, ncols, view):
    _count("lkb_basis_view")
    m = _live(b, MBasis)
    c0, nc = int(col0), int(ncols)
    if c0 < 0 or nc < 1 or c0 + nc > m.data.shape[1]:
        return LKB_ERR_ARG
    return ("__out__", {3: CPtr(MBasis(m.kind, m.data[:, c0:c0 + nc], parent=m))}, LKB_OK)


# ---------------------------------------------------------------------------------------------- operators
def lkb_op_csr_create(interp, ctx, kind, m, n, rowptr, col, val, A):
    _count("lkb_op_csr_create")
    k = KINDS[int(kind)]
    rp, cl, vl = rowptr.obj, col.obj, val.obj
    m, n = int(m), int(n)
    if rp[0] != 0 or np.any(np.diff(rp[:m + 1]) < 0) or (len(cl) and (cl.min() < 0 or cl.max() >= n)):
        return LKB_ERR_ARG
    M = np.zeros((m, n), dtype=DT[k], order="F")
    for i in range(m):
        for q in range(int(rp[i]), int(rp[i + 1])):
            M[i, int(cl[q])] += vl[q]
    return ("__out__", {7: CPtr(MOp(k, M))}, LKB_OK)


def lkb_op_csr_create_dist(interp, ctx, kind, m_global, n_global, row0, m_local, col0, n_local, rowptr, col, val, A):
    """row-sharded CSR: the mock is ONE rank, so the shard must be the whole matrix"""
    _count("lkb_op_csr_create_dist")
    if int(row0) != 0 or int(col0) != 0 or int(m_local) != int(m_global) or int(n_local) != int(n_global):
        return LKB_ERR_ARG
    res = lkb_op_csr_create(interp, ctx, kind, m_global, n_global, rowptr, col, val, A)
    return ("__out__", {11: res[1][7]}, LKB_OK) if isinstance(res, tuple) else res


def _coef(p, n, kind):
    return tuple(DT[kind](v) for v in p.obj.reshape(-1)[:n])


def lkb_op_stencil5_create(interp, ctx, kind, nx, ny, coef5, slow0, nslow_local, A):
    _count("lkb_op_stencil5_create")
    k = KINDS[int(kind)]
    if int(slow0) != 0 or int(nslow_local) != int(ny):
        return LKB_ERR_ARG                      # the mock is one rank
    return ("__out__", {7: CPtr(MOp(k, op=lo.Op.stencil(k, (int(nx), int(ny)), _coef(coef5, 5, k))))}, LKB_OK)


def lkb_op_stencil7_create(interp, ctx, kind, nx, ny, nz, coef7, slow0, nslow_local, A):
    _count("lkb_op_stencil7_create")
    k = KINDS[int(kind)]
    if int(slow0) != 0 or int(nslow_local) != int(nz):
        return LKB_ERR_ARG
    return ("__out__", {8: CPtr(MOp(k, op=lo.Op.stencil(k, (int(nx), int(ny), int(nz)), _coef(coef7, 7, k))))}, LKB_OK)


def lkb_op_destroy(interp, A):
    return LKB_OK


def lkb_op_matvec(interp, A, x, y):
    op, xs, ys = _live_op(A), _live(x, MVec), _live(y, MVec)
    op.n_matvec += 1
    ys.data[...] = op.op.apply(np.ascontiguousarray(xs.data), False)
    return LKB_OK


def lkb_op_rmatvec(interp, A, x, y):
    op, xs, ys = _live_op(A), _live(x, MVec), _live(y, MVec)
    op.n_rmatvec += 1
    ys.data[...] = op.op.apply(np.ascontiguousarray(xs.data), True)
    return LKB_OK


def _live_op(h):
    if not isinstance(h.obj, MOp):
        raise FortranError("mock: not an operator handle")
    return h.obj


# ---------------------------------------------------------------------------------------------- Krylov processes (C oracle)
def _absent_int(v, default):
    return default if int(v) == 0 else int(v)


def _basis_copy(m):
    """the oracle wants a Fortran-contiguous basis: copy in, run, copy back (views of a wider basis are not contiguous in ld)"""
    return np.asfortranarray(m.data.copy())


def lkb_arnoldi(interp, A, X, H, ldh, info, kstart, kend, tol, transpose, blksize):
    _count("lkb_arnoldi")
    op, xb = _live_op(A), _live(X, MBasis)
    p = int(blksize)
    kdim = (xb.data.shape[1] - p) // p
    Hm = _hostmat(H, int(ldh), p * kdim, xb.kind)
    Xc, Hc = _basis_copy(xb), np.asfortranarray(Hm.copy())
    inf = lo.arnoldi(op.op, Xc, Hc, kstart=_absent_int(kstart, 1), kend=_absent_int(kend, kdim),
                     tol=None if tol < 0 else float(tol), trans=bool(int(transpose)), blksize=p)
    xb.data[...] = Xc
    Hm[...] = Hc
    return ("__out__", {4: int(inf)}, LKB_OK)


def lkb_lanczos(interp, A, X, T, ldt, info, kstart, kend, tol):
    _count("lkb_lanczos")
    op, xb = _live_op(A), _live(X, MBasis)
    kdim = xb.data.shape[1] - 1
    Tm = _hostmat(T, int(ldt), kdim, xb.kind)
    Xc, Tc = _basis_copy(xb), np.asfortranarray(Tm.copy())
    inf = lo.lanczos(op.op, Xc, Tc, kstart=_absent_int(kstart, 1), kend=_absent_int(kend, kdim),
                     tol=None if tol < 0 else float(tol))
    xb.data[...] = Xc
    Tm[...] = Tc
    return ("__out__", {4: int(inf)}, LKB_OK)


def lkb_bidiag(interp, A, U, V, B, ldb, info, kstart, kend, tol):
    _count("lkb_bidiag")
    op, ub, vb = _live_op(A), _live(U, MBasis), _live(V, MBasis)
    kdim = ub.data.shape[1] - 1
    Bm = _hostmat(B, int(ldb), kdim, ub.kind)
    Uc, Vc, Bc = _basis_copy(ub), _basis_copy(vb), np.asfortranarray(Bm.copy())
    inf = lo.bidiag(op.op, Uc, Vc, Bc, kstart=_absent_int(kstart, 1), kend=_absent_int(kend, kdim),
                    tol=None if tol < 0 else float(tol))
    ub.data[...], vb.data[...] = Uc, Vc
    Bm[...] = Bc
    return ("__out__", {5: int(inf)}, LKB_OK)


def lkb_qr(interp, Q, col0, p, R, ldr, tol, info):
    _count("lkb_qr")
    qb = _live(Q, MBasis)
    c0, pp = int(col0), int(p)
    Qc = np.asfortranarray(qb.data[:, c0:c0 + pp].copy())
    inf, Rm = lo.qr(Qc, tol=None if tol < 0 else float(tol))
    qb.data[:, c0:c0 + pp] = Qc
    _hostmat(R, int(ldr), pp, qb.kind)[:pp, :] = Rm
    return ("__out__", {6: int(inf)}, LKB_OK)


def lkb_qr_pivoting(interp, Q, col0, p, R, ldr, perm, tol, info):
    _count("lkb_qr_pivoting")
    qb = _live(Q, MBasis)
    c0, pp = int(col0), int(p)
    Qc = np.asfortranarray(qb.data[:, c0:c0 + pp].copy())
    inf, Rm, pm = lo.qr_with_pivoting(Qc, tol=None if tol < 0 else float(tol))
    qb.data[:, c0:c0 + pp] = Qc
    _hostmat(R, int(ldr), pp, qb.kind)[:pp, :] = Rm
    perm.obj.reshape(-1)[:pp] = np.asarray(pm) + 1          # 1-based, as the reference returns it
    return ("__out__", {7: int(inf)}, LKB_OK)


def lkb_dgs_step(interp, X, j, W, wcol0, p, if_chk, beta, ldbeta, info):
    _count("lkb_dgs_step")
    xb, wb = _live(X, MBasis), _live(W, MBasis)
    jj, c0, pp = int(j), int(wcol0), int(p)
    Xc = np.asfortranarray(xb.data[:, :jj].copy())
    Wc = np.asfortranarray(wb.data[:, c0:c0 + pp].copy())
    inf, b = lo.dgs_bas(Wc, Xc, jj)
    wb.data[:, c0:c0 + pp] = Wc
    if beta.obj is not None:
        _hostmat(beta, int(ldbeta), pp, xb.kind)[:jj, :] = b
    return ("__out__", {8: int(inf)}, LKB_OK)


def lkb_orthogonalize_against_basis(interp, X, j, W, wcol0, p, if_chk, beta, ldbeta, info):
    """one pass: beta = X(:j)^H W, W -= X(:j) beta (gram_schmidt.fypp:113-200); info = q when ||W(:, q)|| < atol at entry"""
    _count("lkb_orthogonalize_against_basis")
    xb, wb = _live(X, MBasis), _live(W, MBasis)
    jj, c0, pp = int(j), int(wcol0), int(p)
    Xs, Ws = xb.data[:, :jj], wb.data[:, c0:c0 + pp]
    inf = 0
    for q in range(pp):
        if lo.norm(np.ascontiguousarray(Ws[:, q])) < lo.ATOL[xb.kind]:
            inf = q + 1
    b = np.array([[lo.dot(np.ascontiguousarray(Xs[:, i]), np.ascontiguousarray(Ws[:, q])) for q in range(pp)] for i in range(jj)],
                 dtype=DT[xb.kind]).reshape(jj, pp)
    Ws[...] = Ws - Xs @ b
    if beta.obj is not None:
        _hostmat(beta, int(ldbeta), pp, xb.kind)[:jj, :] = b
    return ("__out__", {8: int(inf)}, LKB_OK)


# ---------------------------------------------------------------------------------------------- solvers (oracle shells)
def _io(p):
    return _deref(p)            # the bind(C) struct is an interpreter object: fields are written in place


def _history(io, res):
    cap = int(io.f["res_cap"])
    buf = io.f["res"].obj
    n = min(len(res), cap)
    if buf is not None:
        buf[:n] = np.asarray(res[:n], dtype=np.float64)
    io.f["res_len"] = len(res)


def _callback(interp, precond, user):
    """the lkb_precond_fn the shim passes (c_funloc of its bind(C) trampoline): called like the library calls it, with the
    'device pointer' of the vector to transform in place; iter / residuals = -1 (absent)"""
    if precond.obj is None:
        return None
    tramp = precond.obj[1]

    def pc(vec, k=None):
        rc, _ = interp.call(tramp, user, CPtr(vec), int(vec.shape[0]), -1 if k is None else int(k), np.float64(-1.0),
                            np.float64(-1.0), CPtr(None))
        if rc != 0:
            raise FortranError("mock: preconditioner callback failed")
    return pc


def _gmres(interp, A, b, x, info, rtol, atol, transpose, io, flexible, precond=None):
    op, bv, xv, st = _live_op(A), _live(b, MVec), _live(x, MVec), _io(io)
    xs = np.ascontiguousarray(xv.data)
    inf, meta = lo.gmres(op.op, np.ascontiguousarray(bv.data), xs, rtol=None if rtol < 0 else float(rtol),
                         atol=None if atol < 0 else float(atol), kdim=int(st.f["kdim"]), maxiter=int(st.f["maxiter"]),
                         trans=bool(int(transpose)), flexible=flexible, precond=precond)
    xv.data[...] = xs
    st.f["n_iter"], st.f["n_inner"], st.f["n_outer"] = meta["n_iter"], meta["n_inner"], meta["n_outer"]
    st.f["converged"], st.f["info"] = int(meta["converged"]), int(inf)
    _history(st, meta["res"])
    return ("__out__", {3: int(inf)}, LKB_OK)


def lkb_gmres(interp, A, b, x, info, rtol, atol, transpose, io):
    _count("lkb_gmres")
    return _gmres(interp, A, b, x, info, rtol, atol, transpose, io, False)


def lkb_gmres_precond(interp, A, b, x, info, rtol, atol, transpose, io, precond, user):
    _count("lkb_gmres_precond")
    return _gmres(interp, A, b, x, info, rtol, atol, transpose, io, False, _callback(interp, precond, user))


def lkb_fgmres(interp, A, b, x, info, rtol, atol, transpose, io, precond, user):
    _count("lkb_fgmres")
    return _gmres(interp, A, b, x, info, rtol, atol, transpose, io, True, _callback(interp, precond, user))


def lkb_vec_wrap(interp, ctx, kind, n_local, n_global, row0, devptr, v):
    """a vector handle over memory the library already owns (the callback's vec_dev)"""
    _count("lkb_vec_wrap")
    k = KINDS[int(kind)]
    arr = devptr.obj
    if not isinstance(arr, np.ndarray) or arr.dtype != DT[k] or arr.shape[0] != int(n_local):
        return LKB_ERR_ARG
    return ("__out__", {6: CPtr(MVec(k, arr))}, LKB_OK)


def lkb_cg_precond(interp, A, b, x, info, rtol, atol, io, precond, user):
    _count("lkb_cg_precond")
    return lkb_cg(interp, A, b, x, info, rtol, atol, io, _precond=_callback(interp, precond, user))


def lkb_cg(interp, A, b, x, info, rtol, atol, io, _precond=None):
    _count("lkb_cg")
    op, bv, xv, st = _live_op(A), _live(b, MVec), _live(x, MVec), _io(io)
    xs = np.ascontiguousarray(xv.data)
    inf, meta = lo.cg(op.op, np.ascontiguousarray(bv.data), xs, rtol=None if rtol < 0 else float(rtol),
                      atol=None if atol < 0 else float(atol), maxiter=int(st.f["maxiter"]), precond=_precond)
    xv.data[...] = xs
    st.f["n_iter"], st.f["converged"], st.f["info"] = meta["n_iter"], int(meta["converged"]), int(inf)
    _history(st, meta["res"])
    return ("__out__", {3: int(inf)}, LKB_OK)


def _opt_vec(h, n, kind, interp):
    if h.obj is None:
        v = interp.rng.standard_normal(n).astype(DT[kind])
        return v
    return np.ascontiguousarray(_live(h, MVec).data)


def lkb_eighs(interp, A, X, nev, eigvals, residuals, info, x0, kdim, tolerance):
    _count("lkb_eighs")
    op, xb = _live_op(A), _live(X, MBasis)
    n, ne = xb.data.shape[0], int(nev)
    ev, res, Xo, k = lo.eighs(op.op, n, ne, _opt_vec(x0, n, xb.kind, interp), kdim=None if int(kdim) <= 0 else int(kdim),
                              tolerance=None if tolerance < 0 else float(tolerance),
                              write_intermediate=bool(Stats.options.get("write_intermediate", 0)))
    xb.data[:, :ne] = Xo
    eigvals.obj[:ne], residuals.obj[:ne] = ev, res
    return ("__out__", {5: int(k)}, LKB_OK)


def lkb_svds(interp, A, U, S, V, nsv, residuals, info, u0, kdim, tolerance):
    _count("lkb_svds")
    op, ub, vb = _live_op(A), _live(U, MBasis), _live(V, MBasis)
    ns = int(nsv)
    sv, res, Uo, Vo, k = lo.svds(op.op, ns, _opt_vec(u0, ub.data.shape[0], ub.kind, interp),
                                 kdim=None if int(kdim) <= 0 else int(kdim),
                                 tolerance=None if tolerance < 0 else float(tolerance),
                                 write_intermediate=bool(Stats.options.get("write_intermediate", 0)))
    ub.data[:, :ns], vb.data[:, :ns] = Uo, Vo
    S.obj[:ns], residuals.obj[:ns] = sv, res
    return ("__out__", {6: int(k)}, LKB_OK)


def lkb_eigs(interp, A, X, nev, eigvals, residuals, info, x0, kdim, tolerance, transpose):
    _count("lkb_eigs")
    op, xb = _live_op(A), _live(X, MBasis)
    n, ne = xb.data.shape[0], int(nev)
    ev, res, Xo, niter = lo.eigs(op.op, n, ne, _opt_vec(x0, n, xb.kind, interp), kdim=None if int(kdim) <= 0 else int(kdim),
                                 tolerance=None if tolerance < 0 else float(tolerance), trans=bool(int(transpose)),
                                 write_intermediate=bool(Stats.options.get("write_intermediate", 0)))
    xb.data[:, :ne] = Xo
    out = eigvals.obj.reshape(-1, order="F")                  # (2, nev) column-major = interleaved (re, im) pairs
    out[0:2 * ne:2], out[1:2 * ne:2] = np.real(ev), np.imag(ev)
    residuals.obj[:ne] = res
    return ("__out__", {5: int(niter)}, LKB_OK)


def lkb_kexpm_vec(interp, c, A, b, tau, tol, info, trans, kdim):
    _count("lkb_kexpm_vec")
    cv, op, bv = _live(c, MVec), _live_op(A), _live(b, MVec)
    out, inf = lo.kexpm_vec(op.op, np.ascontiguousarray(bv.data), float(tau), float(tol), trans=bool(int(trans)),
                            kdim=100 if int(kdim) <= 0 else int(kdim))
    cv.data[...] = out
    return ("__out__", {5: int(inf)}, LKB_OK)


def lkb_kexpm_mat(interp, C, A, B, p, tau, tol, info, trans, kdim):
    _count("lkb_kexpm_mat")
    cb, op, bb = _live(C, MBasis), _live_op(A), _live(B, MBasis)
    pp = int(p)
    out, inf = lo.kexpm_mat(op.op, np.asfortranarray(bb.data[:, :pp].copy()), float(tau), float(tol), trans=bool(int(trans)),
                            kdim=100 if int(kdim) <= 0 else int(kdim))
    cb.data[:, :pp] = out
    return ("__out__", {6: int(inf)}, LKB_OK)


def lkb_krylov_expta(interp, vec_out, A, vec_in, tau, info, trans):
    """kexpm_vec with tol = atol of the kind and kdim = 30 (ExpmLib.fypp:364-392)"""
    _count("lkb_krylov_expta")
    ov, op, iv = _live(vec_out, MVec), _live_op(A), _live(vec_in, MVec)
    out, inf = lo.kexpm_vec(op.op, np.ascontiguousarray(iv.data), float(tau), lo.ATOL[op.kind], trans=bool(int(trans)), kdim=30)
    ov.data[...] = out
    return ("__out__", {4: int(inf)}, LKB_OK)


NATIVES = {name: fn for name, fn in globals().items() if name.startswith("lkb_") and callable(fn)}


def install(interp):
    """register the mock as the library behind the shim's bind(C) interfaces"""
    interp.natives.update(NATIVES)
    Stats.calls, Stats.options = {}, {}
    return Stats
