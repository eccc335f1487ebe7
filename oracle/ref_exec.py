"""ref_exec -- runs procedures of the REFERENCE's own Fortran sources through oracle/f90run.py.  TEST INFRASTRUCTURE ONLY.

Needs /root/reference (or $LK_REFERENCE), so it is used only by tests/golden/make_ref_golden.py (which writes the committed
fixtures tests/golden/ref_*.npz) and by CPU tests that skip when the reference tree is absent (the GPU box).

The concrete vector / operator types are the reference's own test types from src/Utilities/TestUtils.f90 (`vector_rdp`,
`linop_rdp`, `spd_linop_rdp`, `hermitian_linop_cdp`, ...: fixed size `test_size` = 128), i.e. the very types its test-suite
runs arnoldi / lanczos / ... on.  Inputs are written into their `data` components; everything else is the reference's code.
"""
import os

import numpy as np

from . import f90run

REF = os.environ.get("LK_REFERENCE", "/root/reference")
SUFFIX = {"s": "rsp", "d": "rdp", "c": "csp", "z": "cdp"}
DTYPE = {"s": np.float32, "d": np.float64, "c": np.complex64, "z": np.complex128}

FILES = [
    "src/Constants.f90", "src/AbstractTypes/AbstractVectors.f90", "src/AbstractTypes/AbstractLinops.f90",
    "src/AbstractTypes/AbstractSystems.f90", "src/Krylov/BaseKrylov.f90", "src/Krylov/utilities.f90",
    "src/Krylov/gram_schmidt.f90", "src/Krylov/qr.f90", "src/Krylov/arnoldi.f90", "src/Krylov/lanczos.f90",
    "src/Krylov/golub_kahan.f90", "src/Utilities/Utils.f90", "src/Utilities/submodule_utility_functions.f90",
    "src/Utilities/TestUtils.f90", "src/IterativeSolvers/IterativeSolvers.f90", "src/IterativeSolvers/CG/CG.f90",
    "src/IterativeSolvers/GMRES/gmres.f90", "src/IterativeSolvers/GMRES/fgmres.f90",
    "src/IterativeSolvers/EIGHS/eighs.f90", "src/IterativeSolvers/SVDS/svd_solvers.f90", "src/Expm/ExpmLib.f90",
]


def available() -> bool:
    return os.path.isfile(os.path.join(REF, FILES[0]))


_interp = None


def interp() -> "f90run.Interp":
    """the reference, parsed once per process"""
    global _interp
    if _interp is None:
        prog = f90run.Program()
        for f in FILES:
            prog.load(os.path.join(REF, f))
        _interp = f90run.Interp(prog)
        _interp.rng = np.random.default_rng(1000)
    return _interp


def test_size() -> int:
    return int(interp().p.globals["test_size"])


# ---------------------------------------------------------------------------------------------- inputs (bit-reproducible)
def pseudo(shape, seed, kind="d"):
    """Deterministic pseudo-random numbers in [-0.5, 0.5) from integer hashing only (exact in every precision, so the test that
    re-creates the inputs of a committed fixture gets them bit for bit, on any machine and numpy version)."""
    n = int(np.prod(shape))

    def one(sd):
        x = (np.arange(n, dtype=np.uint64) + np.uint64(sd) * np.uint64(0x9E3779B1)) & np.uint64(0xFFFFFFFF)
        x = (x * np.uint64(2654435761)) & np.uint64(0xFFFFFFFF)
        x ^= x >> np.uint64(16)
        x = (x * np.uint64(2246822519)) & np.uint64(0xFFFFFFFF)
        x ^= x >> np.uint64(13)
        x = (x * np.uint64(3266489917)) & np.uint64(0xFFFFFFFF)
        x ^= x >> np.uint64(16)
        return ((x >> np.uint64(8)).astype(np.float64) / float(1 << 24) - 0.5)          # 24 bits: exact in fp32 too
    v = one(seed)
    if kind in "cz":
        v = v + 1j * one(seed + 7919)
    return np.asfortranarray(v.astype(DTYPE[kind]).reshape(shape, order="F"))


# ---------------------------------------------------------------------------------------------- reference objects
def vector(kind, data=None):
    it = interp()
    v = it.new_inst("vector_" + SUFFIX[kind])
    if data is not None:
        v.f["data"][...] = data
    return v


def basis(kind, ncols, first=None):
    X = np.empty(ncols, dtype=object)
    for i in range(ncols):
        X[i] = vector(kind)
    if first is not None:
        first = np.asarray(first)
        if first.ndim == 1:
            first = first[:, None]
        for i in range(first.shape[1]):
            X[i].f["data"][...] = first[:, i]
    return X


def basis_data(X):
    return np.asfortranarray(np.column_stack([v.f["data"] for v in X]))


def linop(kind, A, sym=False):
    it = interp()
    name = ("spd_linop_" if kind in "sd" else "hermitian_linop_") + SUFFIX[kind] if sym else "linop_" + SUFFIX[kind]
    op = it.new_inst(name)
    op.f["data"][...] = A
    return op


def call(name, *args, **kwargs):
    """-> (function result, {position / keyword: final value}) ; arrays and objects are updated in place"""
    return interp().call(name, *args, **kwargs)
