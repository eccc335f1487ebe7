"""Golden fixtures (tests/golden/*.npz, produced by tests/golden/make_golden.py from the pinned oracle):
CPU tests check that the oracle still reproduces them; GPU tests compare the CUDA path with the
committed numbers through the C ABI."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
sys.path.insert(0, GOLD)


def _load(name):
    return np.load(os.path.join(GOLD, name))


def _tol(kind):
    return 1e-10 if kind in "dz" else 1e-4


def _rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


@pytest.mark.parametrize("kind", list("sdcz"))
def test_oracle_reproduces_config1(oracle, kind):
    import make_golden
    g = _load(f"config1_{kind}.npz")
    r = make_golden.config1(kind)
    assert r["info"] == int(g["info"]) == 0
    assert _rel(r["H"], g["H"]) < (1e-13 if kind in "dz" else 1e-5)      # same machine arithmetic up to FMA contraction
    assert np.array_equal(r["A"], g["A"]) and np.array_equal(r["x0"], g["x0"])


@pytest.mark.parametrize("kind", list("dz"))
def test_oracle_reproduces_poisson(oracle, kind):
    import make_golden
    g = _load(f"poisson_{kind}.npz")
    r = make_golden.poisson(kind)
    assert _rel(r["H"], g["H"]) < 1e-12 and _rel(r["T"], g["T"]) < 1e-12
    assert r["gmres_info"] == int(g["gmres_info"]) and _rel(r["gmres_x"], g["gmres_x"]) < 1e-9
    assert np.array_equal(r["x0_head"], g["x0_head"])                   # counter RNG is bit-stable


@pytest.mark.gpu
@pytest.mark.parametrize("kind", list("sdcz"))
def test_gpu_matches_golden_config1(kind):
    import lightkrylov_b200 as lk
    g = _load(f"config1_{kind}.npz")
    ctx = lk.Context(0)
    n, kdim = 128, 64
    A = lk.LinOp.dense(ctx, g["A"])
    X = lk.Basis(ctx, kind, n, kdim + 1).put(g["x0"])
    H = np.zeros((kdim + 1, kdim), dtype=lk.DTYPES[kind], order="F")
    assert lk.arnoldi(A, X, H) == int(g["info"])
    assert _rel(H, g["H"]) < _tol(kind)
    assert _rel(X.get(kdim, 1)[:, 0], g["X_last"]) < _tol(kind)
    del X, A
    ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("kind", list("dz"))
def test_gpu_matches_golden_poisson(kind):
    import lightkrylov_b200 as lk
    g = _load(f"poisson_{kind}.npz")
    nx, ny, kdim = (int(v) for v in g["dims"])
    n = nx * ny
    ctx = lk.Context(0)
    A = lk.LinOp.stencil5(ctx, kind, nx, ny, (4.0, -1.0, -1.0, -1.0, -1.0))
    X = lk.Basis(ctx, kind, n, kdim + 1)
    x0 = X.col(0).fill_random("uniform", 42); x0.scal(1.0 / x0.norm())
    np.testing.assert_allclose(x0.get()[:8], g["x0_head"], rtol=1e-13)
    H = np.zeros((kdim + 1, kdim), dtype=lk.DTYPES[kind], order="F")
    assert lk.arnoldi(A, X, H) == int(g["info"]) and _rel(H, g["H"]) < 1e-10
    Xl = lk.Basis(ctx, kind, n, kdim + 1); xl = Xl.col(0).fill_random("uniform", 42); xl.scal(1.0 / xl.norm())
    T = np.zeros((kdim + 1, kdim), dtype=lk.DTYPES[kind], order="F")
    assert lk.lanczos(A, Xl, T) == int(g["linfo"]) and _rel(T, g["T"]) < 1e-10
    A2 = lk.LinOp.stencil5(ctx, kind, nx, ny, (6.0, -1.3, -0.7, -1.2, -0.8))
    b = lk.Vector(ctx, kind, n).fill_random("uniform", 43); x = lk.Vector(ctx, kind, n)
    info, meta = lk.gmres(A2, b, x, kdim=20, maxiter=20)
    assert info == int(g["gmres_info"]) and _rel(x.get(), g["gmres_x"]) < 1e-10
    assert _rel(np.array(meta["res"]), g["gmres_res"]) < 1e-10
    del X, Xl, A, A2, b, x, x0, xl
    ctx.close()
