"""Shared fixtures/utilities of the GPU parity tests (seeded inputs identical on oracle and device)."""
import numpy as np

POISSON5 = (4.0, -1.0, -1.0, -1.0, -1.0)
CONVDIFF7 = (6.0, -1.3, -0.7, -1.2, -0.8, -1.1, -0.9)          # config C3: -1 -/+ gamma_d
LAPLACE7 = (6.0, -1.0, -1.0, -1.0, -1.0, -1.0, -1.0)           # config C4


def randn(rng, shape, dtype):
    a = rng.standard_normal(shape)
    if np.issubdtype(np.dtype(dtype), np.complexfloating):
        a = a + 1j * rng.standard_normal(shape)
    return np.asfortranarray(a.astype(dtype))


def rel_normwise(a, b):
    """max|a-b| / max|b|  (SURVEY hard parts: entrywise relative parity is impossible on noise entries)."""
    den = np.abs(b).max()
    return float(np.abs(a - b).max() / (den if den > 0 else 1.0))


def tol_for(kind):
    return 1e-10 if kind in ("d", "z") else 1e-4                # BASELINE.json north_star


def orth_tol(kind):
    return 1e-12 if kind in ("d", "z") else 2e-5


def random_csr(rng, m, n, per_row, dtype):
    import scipy.sparse as sp
    cols = np.sort(rng.integers(0, n, size=(m, per_row)), axis=1)
    vals = randn(rng, (m, per_row), dtype)
    rows = np.repeat(np.arange(m), per_row)
    S = sp.coo_matrix((np.asarray(vals).ravel(), (rows, cols.ravel())), shape=(m, n)).tocsr()  # duplicates summed
    S.sort_indices()
    return S
