#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ from the CPU oracle.

The reference (Fortran) ships no golden vectors and cannot be built in this image, so these
fixtures are produced by the oracle AFTER it has been pinned by the reference's own property /
known-answer assertions (tests/test_oracle_pins.py).  They freeze the oracle's outputs on seeded
inputs (BASELINE config 1 and small stencil cases) so that (a) the oracle itself is regression-
tested on CPU and (b) the CUDA path is compared against committed numbers, not only against a
live oracle run.

    python tests/golden/make_golden.py        # rewrites tests/golden/*.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import lk_oracle as lo  # noqa: E402

lo.set_threads(1)          # sequential summation order: the reference is serial


def randn(rng, shape, dtype):
    a = rng.standard_normal(shape)
    if np.issubdtype(np.dtype(dtype), np.complexfloating):
        a = a + 1j * rng.standard_normal(shape)
    return np.asfortranarray(a.astype(dtype))


def config1(kind):
    """BASELINE configs[0]: arnoldi kdim=64 on a random dense linop, n=128 (A seed 1, x0 seed 2)."""
    dt = lo.DTYPES[kind]; n, kdim = 128, 64
    A = randn(np.random.default_rng(1), (n, n), dt)
    x0 = randn(np.random.default_rng(2), n, dt); lo.normalize(x0)
    X = np.zeros((n, kdim + 1), dtype=dt, order="F"); X[:, 0] = x0
    H = np.zeros((kdim + 1, kdim), dtype=dt, order="F")
    info = lo.arnoldi(lo.Op.dense(A), X, H)
    return dict(A=A, x0=x0, H=H, X_last=X[:, -1].copy(), info=info)


def poisson(kind, nx=48, ny=40, kdim=24):
    """Config-2 operator (5-point Poisson) on a small grid, x0 = U(0,1) seed 42."""
    dt = lo.DTYPES[kind]; n = nx * ny
    x0 = lo.fill(n, kind, "uniform", 42); lo.normalize(x0)
    X = np.zeros((n, kdim + 1), dtype=dt, order="F"); X[:, 0] = x0
    H = np.zeros((kdim + 1, kdim), dtype=dt, order="F")
    info = lo.arnoldi(lo.Op.stencil(kind, (nx, ny), (4.0, -1.0, -1.0, -1.0, -1.0)), X, H)
    T = np.zeros((kdim + 1, kdim), dtype=dt, order="F")
    Xl = np.zeros((n, kdim + 1), dtype=dt, order="F"); Xl[:, 0] = x0
    linfo = lo.lanczos(lo.Op.stencil(kind, (nx, ny), (4.0, -1.0, -1.0, -1.0, -1.0)), Xl, T)
    b = lo.fill(n, kind, "uniform", 43)
    x = np.zeros(n, dtype=dt)
    ginfo, gmeta = lo.gmres(lo.Op.stencil(kind, (nx, ny), (6.0, -1.3, -0.7, -1.2, -0.8)), b, x, kdim=20, maxiter=20)
    return dict(dims=np.array([nx, ny, kdim]), H=H, info=info, T=T, linfo=linfo, gmres_x=x, gmres_info=ginfo,
                gmres_res=np.array(gmeta["res"]), x0_head=x0[:8].copy())


if __name__ == "__main__":
    for kind in "sdcz":
        np.savez_compressed(os.path.join(HERE, f"config1_{kind}.npz"), **config1(kind))
    for kind in "dz":
        np.savez_compressed(os.path.join(HERE, f"poisson_{kind}.npz"), **poisson(kind))
    print("golden fixtures written to", HERE)
