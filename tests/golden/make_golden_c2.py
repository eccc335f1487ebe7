#!/usr/bin/env python
"""Golden Hessenberg matrix of BASELINE configs[1] at FULL size (5-point Poisson 4096^2, n = 16.8M, fp64,
arnoldi kdim = 128, x0 = U[0,1) seed 42 normalised), produced by the CPU oracle (all 128 steps, the
reference's per-vector op sequence).  129 x 128 doubles = 132 KB.

    python tests/golden/make_golden_c2.py [threads]      # ~5-10 min, 17.3 GB of host memory

The GPU parity test (tests/test_gpu_parity.py::test_arnoldi_full_size_c2) and bench.py's `parity` record
(every N in {1,2,4,8}) compare the CUDA path's H and Ritz values with this matrix at 1e-10.
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import lk_oracle as lo  # noqa: E402

if __name__ == "__main__":
    threads = int(sys.argv[1]) if len(sys.argv) > 1 else lo.max_threads()
    lo.set_threads(threads)
    nx = ny = 4096; n = nx * ny; kdim = 128
    X = np.zeros((n, kdim + 1), order="F")
    X[:, 0] = lo.fill(n, "d", "uniform", 42); lo.normalize(X[:, 0])
    H = np.zeros((kdim + 1, kdim), order="F")
    t0 = time.time()
    info = lo.arnoldi(lo.Op.stencil("d", (nx, ny), (4.0, -1.0, -1.0, -1.0, -1.0)), X, H)
    dt = time.time() - t0
    ritz = np.sort(np.linalg.eigvals(H[:kdim, :kdim]).real)
    # orthonormality of the oracle's own basis (the reference's test criterion), first/last 8 columns
    G = X[:, :8].T @ X[:, -8:]
    np.savez_compressed(os.path.join(HERE, "c2_full_H.npz"), H=H, info=info, ritz=ritz, threads=threads,
                        seconds=dt, x0_head=X[:8, 0].copy(), xlast_head=X[:8, kdim].copy(), cross_gram_max=np.abs(G).max())
    print("C2 golden written: info", info, "seconds", dt, "max|G_cross|", np.abs(G).max())
