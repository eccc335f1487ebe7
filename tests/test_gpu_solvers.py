"""GPU parity of the solver shells built on the device Krylov step: eigs (+ Krylov-Schur),
eighs, svds -- the reference's known-answer tests (test/TestIterativeSolvers.fypp) run through
the C ABI, plus entrywise comparison with the numpy/scipy oracle shells on identical inputs."""
import numpy as np
import pytest

from helpers import CONVDIFF7, LAPLACE7, POISSON5, randn, random_csr, rel_normwise

pytestmark = pytest.mark.gpu
N = 128


@pytest.fixture(scope="module")
def lk():
    import lightkrylov_b200 as lk
    lk.set_lapack_from_scipy()
    return lk


@pytest.fixture(scope="module")
def ctx(lk):
    c = lk.Context(0)
    yield c
    c.close()


def _toeplitz(n, sub, diag, sup, dtype=np.float64):
    A = np.zeros((n, n), dtype=dtype); i = np.arange(n)
    A[i, i] = diag; A[i[1:], i[:-1]] = sub; A[i[:-1], i[1:]] = sup
    return np.asfortranarray(A)


def _match(a, b):
    return np.abs(a[:, None] - b[None, :]).min(axis=1).max()


def test_eigs_known_answer_full(lk, ctx, oracle):
    """TestIterativeSolvers.fypp:134-197 (nev = n = 128, kdim at its default 4*nev = 512 > n as in the reference's test: the
    literal post-convergence krylov_schur then retains all n Ritz values, see tests/test_oracle_pins.py)."""
    a, b = 1.0, 0.5
    Ah = _toeplitz(N, -b, a, b)
    x0h = np.random.default_rng(20).standard_normal(N)
    A = lk.LinOp.dense(ctx, Ah); X = lk.Basis(ctx, "d", N, N)
    x0 = lk.Vector(ctx, "d", N).put(x0h)
    ev, res, info = lk.eigs(A, X, N, x0=x0)
    true = a + 2j * b * np.cos(np.arange(1, N + 1) * np.pi / (N + 1))
    assert _match(ev, true) < lk.RTOL["d"] and _match(true, ev) < lk.RTOL["d"]
    evo, reso, Xo, infoo = oracle.eigs(oracle.Op.dense(Ah), N, N, x0h)
    assert info == infoo
    assert _match(ev, evo) < 1e-10


@pytest.mark.parametrize("kind", ["d", "z", "s"])
def test_eigs_krylov_schur_vs_oracle(lk, ctx, oracle, kind):
    """TestIterativeSolvers.fypp:161-209: nev = 8, kdim = 32, restarts; same niter, Ritz values to 1e-10."""
    dt = lk.DTYPES[kind]; a, b, nev = 1.0, 0.5, 8
    Ah = _toeplitz(N, -b, a, b, dt)
    x0h = randn(np.random.default_rng(21), N, dt)
    A = lk.LinOp.dense(ctx, Ah); X = lk.Basis(ctx, kind, N, nev)
    ev, res, info = lk.eigs(A, X, nev, x0=lk.Vector(ctx, kind, N).put(x0h), kdim=4 * nev)
    evo, reso, Xo, infoo = oracle.eigs(oracle.Op.dense(Ah), N, nev, x0h, kdim=4 * nev)
    tol = 1e-10 if kind in "dz" else 1e-4
    if kind in "dz":
        assert info == infoo
    assert _match(ev, evo) < tol * np.abs(evo).max()
    true = a + 2j * b * np.cos(np.arange(1, N + 1) * np.pi / (N + 1))
    lead = true[np.argsort(-np.abs(true))][:nev]
    assert _match(ev, lead) < (1e-6 if kind in "dz" else 1e-3)
    # eigenvectors: A v = lambda v through the real-pair convention / directly for complex kinds
    Xg = X.get().astype(np.complex128)
    Ac = Ah.astype(np.complex128)
    i = 0
    while i < nev:
        if kind in "cz" or ev[i].imag == 0:
            v, lam = Xg[:, i], ev[i]; i += 1
        elif i + 1 < nev:
            v = Xg[:, i] + 1j * Xg[:, i + 1] if ev[i].imag > 0 else Xg[:, i + 1] + 1j * Xg[:, i]
            lam = ev[i] if ev[i].imag > 0 else ev[i + 1]; i += 2
        else:
            break
        # fp32: the solver stops at Ritz residual < rtol_sp = 1e-3, so eigenpairs are only that accurate
        assert np.linalg.norm(Ac @ v - lam * v) < (1e-6 if kind in "dz" else 1e-2) * np.linalg.norm(v)


def test_krylov_schur_restart_relation(lk, ctx, oracle):
    """TestKrylov.fypp:298-347 on the device: A X_n = X_{n+1} H(:n+1,:n), orthonormal; vs oracle."""
    rng = np.random.default_rng(22); kdim = 32
    Ah = randn(rng, (N, N), np.float64) / np.sqrt(N)
    x0 = randn(rng, N, np.float64); oracle.normalize(x0)
    A = lk.LinOp.dense(ctx, Ah)
    X = lk.Basis(ctx, "d", N, kdim + 1).put(x0); H = np.zeros((kdim + 1, kdim), order="F")
    assert lk.arnoldi(A, X, H) == 0
    nk = lk.krylov_schur(X, H, kdim)
    Xo = np.zeros((N, kdim + 1), order="F"); Xo[:, 0] = x0; Ho = np.zeros_like(H)
    oracle.arnoldi(oracle.Op.dense(Ah), Xo, Ho); nko = oracle.krylov_schur(Xo, Ho)
    assert nk == nko and 0 < nk < kdim
    Xg = X.get()
    assert np.abs(Ah @ Xg[:, :nk] - Xg[:, :nk + 1] @ H[:nk + 1, :nk]).max() < lk.RTOL["d"]
    assert np.abs(Xg[:, :nk + 1].T @ Xg[:, :nk + 1] - np.eye(nk + 1)).max() < 1e-12
    assert not Xg[:, nk + 1:].any() and not H[nk + 1:, :].any() and not H[:, nk:].any()
    # the reordered real Schur form is not unique (2x2 block standardisation, swap order), so compare the
    # invariants: the retained Ritz values and the residual row norm
    assert _match(np.linalg.eigvals(H[:nk, :nk]), np.linalg.eigvals(Ho[:nk, :nk])) < 1e-10
    assert abs(np.linalg.norm(H[nk, :nk]) - np.linalg.norm(Ho[nk, :nk])) < 1e-10
    # resume the factorisation from the restarted state (the way eigs does)
    assert lk.arnoldi(A, X, H, kstart=nk + 1, kend=kdim) == 0
    Xg = X.get()
    assert np.abs(Ah @ Xg[:, :kdim] - Xg @ H).max() < lk.RTOL["d"]


@pytest.mark.parametrize("kind", ["d", "s", "z"])
def test_eighs_known_answer_and_oracle(lk, ctx, oracle, kind):
    """TestIterativeSolvers.fypp:254-307."""
    dt = lk.DTYPES[kind]; a, b, nev = 2.0, -1.0, 8
    Ah = _toeplitz(N, b, a, b, dt)
    x0h = randn(np.random.default_rng(23), N, dt)
    A = lk.LinOp.dense(ctx, Ah); X = lk.Basis(ctx, kind, N, nev)
    ev, res, info = lk.eighs(A, X, nev, x0=lk.Vector(ctx, kind, N).put(x0h), kdim=N)
    true = a + 2 * abs(b) * np.cos(np.arange(1, N + 1) * np.pi / (N + 1))
    # fp32 stops as soon as ANY nev Ritz residuals are < rtol_sp = 1e-3 (eighs.fypp:94-99): loose known answer
    assert np.abs(ev - true[:nev]).max() < (lk.RTOL[kind] * 4 if kind in "dz" else 3e-2)
    evo, reso, Xo, infoo = oracle.eighs(oracle.Op.dense(Ah), N, nev, x0h, kdim=N)
    if kind in "dz":
        assert info == infoo
    assert np.abs(ev - evo).max() < (1e-10 if kind in "dz" else 1e-4) * np.abs(evo).max()
    Xg = X.get()
    assert np.abs(Ah @ Xg - Xg * ev.astype(dt)).max() < (1e-6 if kind in "dz" else 5e-2)
    assert np.abs(Xg.conj().T @ Xg - np.eye(nev)).max() < lk.RTOL[kind] * 4


@pytest.mark.parametrize("kind", ["d", "z"])
def test_svds_known_answer_and_oracle(lk, ctx, oracle, kind):
    """TestIterativeSolvers.fypp:440-489 (Strang matrix) + a rectangular random CSR vs oracle."""
    dt = lk.DTYPES[kind]; nsv = 8
    Ah = _toeplitz(N, -1.0, 2.0, -1.0, dt)
    u0h = randn(np.random.default_rng(24), N, dt)
    A = lk.LinOp.dense(ctx, Ah)
    U = lk.Basis(ctx, kind, N, nsv); V = lk.Basis(ctx, kind, N, nsv)
    S, res, info = lk.svds(A, U, V, nsv, u0=lk.Vector(ctx, kind, N).put(u0h), kdim=N)
    true = 2 * (1 + np.cos(np.arange(1, N + 1) * np.pi / (N + 1)))
    assert np.abs(S - true[:nsv]).max() < lk.RTOL[kind]
    Ug, Vg = U.get(), V.get()
    assert np.abs(Ah @ Vg - Ug * S.astype(dt)).max() < 1e-6
    assert np.abs(Ug.conj().T @ Ug - np.eye(nsv)).max() < lk.RTOL[kind]
    So, reso, Uo, Vo, infoo = oracle.svds(oracle.Op.dense(Ah), nsv, u0h, kdim=N)
    assert info == infoo and np.abs(S - So).max() < 1e-10 * So.max()
    # config-5 shaped: rectangular CSR, cdp/rdp
    m, n = 900, 700
    Sp = random_csr(np.random.default_rng(46), m, n, 32, dt)
    Ac = lk.LinOp.csr(ctx, m, n, Sp.indptr, Sp.indices, Sp.data.astype(dt))
    Aco = oracle.Op.csr(m, n, Sp.indptr, Sp.indices, Sp.data.astype(dt))
    u0h = oracle.fill(m, kind, "normal", 47)
    U = lk.Basis(ctx, kind, m, nsv); V = lk.Basis(ctx, kind, n, nsv)
    S, res, info = lk.svds(Ac, U, V, nsv, u0=lk.Vector(ctx, kind, m).put(u0h), kdim=32)
    So, reso, Uo, Vo, infoo = oracle.svds(Aco, nsv, u0h, kdim=32)
    assert info == infoo and np.abs(S - So).max() < 1e-10 * So.max()
    sv_true = np.linalg.svd(Sp.toarray(), compute_uv=False)
    assert abs(S[0] - sv_true[0]) < 1e-6 * sv_true[0]


def test_eigs_convdiff_stencil_vs_oracle(lk, ctx, oracle):
    """Config-3 shaped: nonsymmetric 7-point convection-diffusion, eigs nev=4 with Krylov-Schur."""
    dims = (12, 10, 8); n = int(np.prod(dims)); nev = 4
    A = lk.LinOp.stencil7(ctx, "d", *dims, CONVDIFF7); Ao = oracle.Op.stencil("d", dims, CONVDIFF7)
    x0h = oracle.fill(n, "d", "uniform", 44)
    X = lk.Basis(ctx, "d", n, nev)
    ev, res, info = lk.eigs(A, X, nev, x0=lk.Vector(ctx, "d", n).put(x0h), kdim=24, tolerance=1e-8)
    evo, reso, Xo, infoo = oracle.eigs(Ao, n, nev, x0h, kdim=24, tolerance=1e-8)
    assert info == infoo
    assert _match(ev, evo) < 1e-10 * np.abs(evo).max()


@pytest.mark.parametrize("kind", ["d", "z", "s"])
def test_kexpm_vec_vs_oracle(lk, ctx, oracle, kind):
    """kexpm_vec (src/Expm/ExpmLib.fypp:128-232): c = exp(tau A) b on the device vs the oracle restatement:
    same `info` (Krylov dimension used) and the same vector."""
    dims = (24, 20); n = 480
    coef = (-4.0, 1.0, 1.0, 1.0, 1.0)
    A = lk.LinOp.stencil5(ctx, kind, *dims, coef); Ao = oracle.Op.stencil(kind, dims, coef)
    bh = oracle.fill(n, kind, "uniform", 3)
    b = lk.Vector(ctx, kind, n).put(bh); c = lk.Vector(ctx, kind, n)
    tol = 1e-10 if kind in "dz" else 1e-5
    info = lk.kexpm(c, A, b, 0.1, tol)
    co, oinfo = oracle.kexpm_vec(Ao, bh, 0.1, tol)
    assert info == oinfo and info > 1
    assert np.linalg.norm(c.get() - co) < (1e-10 if kind in "dz" else 1e-4) * np.linalg.norm(co)
    # zero input => zero output (ExpmLib.fypp:180-184); transpose of the symmetric operator gives the same vector
    z = lk.Vector(ctx, kind, n); c2 = lk.Vector(ctx, kind, n).put(bh)
    lk.kexpm(c2, A, z, 0.1, tol)
    assert not c2.get().any()
    ct = lk.Vector(ctx, kind, n)
    assert lk.kexpm(ct, A, b, 0.1, tol, trans=True) == info
    assert np.linalg.norm(ct.get() - co) < (1e-10 if kind in "dz" else 1e-4) * np.linalg.norm(co)


@pytest.mark.parametrize("kind", ["d", "z"])
def test_kexpm_vec_breakdown_literal_info(lk, ctx, oracle, kind):
    """Arnoldi breakdown inside kexpm_vec: b spans a 3-dimensional invariant subspace of a diagonal operator.  The reference
    overwrites `info` with -2 before its merge(0, ..., info == k), so the error estimate is not zeroed at the breakdown step;
    the loop continues on the refilled vector and stops two steps later with info = kp = 5 and the exact vector
    (ExpmLib.fypp:199-226; literal flow pinned on the CPU by tests/test_oracle_second_opinion.py)."""
    dt = lk.DTYPES[kind]; n = 96
    D = np.asfortranarray(np.diag(np.arange(1, n + 1)).astype(dt))
    A = lk.LinOp.dense(ctx, D); Ao = oracle.Op.dense(D)
    bh = np.zeros(n, dtype=dt); bh[:3] = [1.0, 2.0, -1.5]
    b = lk.Vector(ctx, kind, n).put(bh); c = lk.Vector(ctx, kind, n)
    info = lk.kexpm(c, A, b, 0.2, 1e-12)
    co, oinfo = oracle.kexpm_vec(Ao, bh, 0.2, 1e-12)
    assert info == oinfo == 5
    exact = np.exp(0.2 * np.arange(1, n + 1)) * bh
    assert np.linalg.norm(c.get() - exact) < 1e-12 * np.linalg.norm(exact)
