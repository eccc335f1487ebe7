"""The k x k host algebra of the solver shells (eig / schur + ordschur / eigh / svd) runs LAPACK IN THE PRECISION OF THE KIND,
as the reference does through stdlib (sgeev ... for rsp / csp, dgeev ... for rdp / cdp).  These helpers need no GPU, so they
are exercised here directly: a small C++ harness includes lkb_eig.cu, calls them on random 24 x 24 matrices through the same
dlopen'd provider the product uses (scipy's OpenBLAS) and prints the residuals of A V = V L, A Z = Z T, A V = U S."""
import glob
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "lightkrylov_b200", "csrc")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


@pytest.mark.skipif(not os.path.exists(NVCC) and shutil.which("nvcc") is None, reason="needs nvcc to compile the harness")
def test_host_lapack_layer_all_four_kinds(tmp_path):
    import scipy
    import __graft_entry__
    __graft_entry__.build()
    exe = str(tmp_path / "host_lapack_harness")
    nvcc = NVCC if os.path.exists(NVCC) else shutil.which("nvcc")
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O1", "-std=c++17", "-I", CSRC,
                           os.path.join(ROOT, "tests", "host_lapack_harness.cu"), "-o", exe, "-L", CSRC, "-llkb", "-ldl",
                           "-Xlinker", "-rpath=" + CSRC])
    blas = glob.glob(os.path.join(os.path.dirname(os.path.dirname(scipy.__file__)), "scipy.libs", "libscipy_openblas*.so"))[0]
    out = subprocess.run([exe, blas], check=True, capture_output=True, text=True).stdout
    lines = [ln.split() for ln in out.strip().splitlines()]
    assert lines[0] == ["sortidx", "1", "2", "6", "3", "0", "4", "5"], out     # stable non-increasing order (stdlib sort_index)
    lines = lines[1:]
    assert len(lines) == 16, out
    seen = set()
    for ln in lines:
        kind, what, r = ln[0], ln[1], float(ln[3])
        seen.add((kind, what))
        assert r < (1e-12 if kind in "dz" else 5e-5), out                 # residual in the precision of the kind
        if kind in "sc":
            assert r > 1e-9, "fp32 kinds must run the s / c LAPACK routines, not be promoted to double: " + out
        if what == "schur":
            assert 8 <= int(ln[5]) <= 14, out                             # median selector keeps about half
        if what in ("eigh", "svd"):
            assert ln[5] == "1", out                                      # ascending eigenvalues / descending singular values
    assert seen == {(k, w) for k in "dzsc" for w in ("eig", "schur", "eigh", "svd")}


def _build_harness(name, tmp_path):
    import __graft_entry__
    __graft_entry__.build()
    exe = str(tmp_path / name)
    nvcc = NVCC if os.path.exists(NVCC) else shutil.which("nvcc")
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O1", "-std=c++17", "-I", CSRC,
                           os.path.join(ROOT, "tests", name + ".cu"), "-o", exe, "-L", CSRC, "-llkb", "-ldl",
                           "-Xlinker", "-rpath=" + CSRC], stderr=subprocess.DEVNULL)
    return exe


@pytest.mark.skipif(not os.path.exists(NVCC) and shutil.which("nvcc") is None, reason="needs nvcc to compile the harness")
def test_host_dense_expm_and_formatter(tmp_path):
    """The product's restatement of stdlib's `expm` (Pade 10 + scaling and squaring, lkb_expm.cu) without a GPU: the reference's
    own dense test (test/TestExpmlib.fypp `test_dense_expm_*`: nilpotent shift, E(i, i+j) = m^j / j!), exp(A) exp(-A) = I,
    unitarity for a skew-Hermitian argument, and entry-by-entry agreement with scipy.linalg.expm on a random complex matrix;
    plus the Fortran E16.9 formatter of write_results."""
    import numpy as np
    import scipy.linalg as sla
    exe = _build_harness("host_expm_harness", tmp_path)
    rng = np.random.default_rng(8)
    n = 23
    A = (rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))) * 0.7
    text = str(n) + "\n" + "\n".join(f"{float(v.real)!r} {float(v.imag)!r}" for v in A.ravel(order="F")) + "\n"
    out = subprocess.run([exe], input=text, check=True, capture_output=True, text=True).stdout.splitlines()
    kv = {ln.split()[0]: ln.split()[1:] for ln in out if ln.split()[0] in ("nilpotent", "inverse", "unitary")}
    assert float(kv["nilpotent"][0]) < 1e-12
    assert float(kv["inverse"][0]) < 1e-10 * float(kv["inverse"][2]) ** 2      # relative to ||exp(A)|| ||exp(-A)||
    assert float(kv["unitary"][0]) < 1e-12
    i0 = out.index(f"matrix {n}")
    E = np.array([[float(x) for x in ln.split()] for ln in out[i0 + 1:i0 + 1 + n * n]])
    E = (E[:, 0] + 1j * E[:, 1]).reshape((n, n), order="F")
    ref = sla.expm(A)
    assert np.abs(E - ref).max() < 1e-12 * np.abs(ref).max()
    fmt = [ln[5:-1] for ln in out if ln.startswith("fmt [")]
    assert fmt == [" 0.000000000E+00", " 0.100000000E+01", "-0.100000000E+01", " 0.100000000E+00", " 0.123456789E+06",
                   "-0.100000000E-03", " 0.100000000+101", " 0.300000000-309", " 0.100000000E+01"]
    assert all(len(f) == 16 for f in fmt)


@pytest.mark.skipif(not os.path.exists(NVCC) and shutil.which("nvcc") is None, reason="needs nvcc to compile the harness")
def test_host_givens_helpers(tmp_path):
    """lkb_solvers.cu's host restatements behind apply_givens_rotation (submodule_utility_functions.fypp:169-204): `lartg`
    against the provider's LAPACK dlartg on 211 (f, g) pairs incl. zeros and extreme magnitudes; the real form is an
    orthogonal rotation that annihilates h(k+1); the complex form is the reference's literal one."""
    import scipy
    exe = _build_harness("host_givens_harness", tmp_path)
    blas = glob.glob(os.path.join(os.path.dirname(os.path.dirname(scipy.__file__)), "scipy.libs", "libscipy_openblas*.so"))[0]
    out = subprocess.run([exe, blas], check=True, capture_output=True, text=True).stdout.splitlines()
    vals = {ln.split()[0]: [float(x) for x in ln.split()[1:] if x[0].isdigit()] for ln in out}
    assert vals["lartg"][0] < 1e-14
    assert vals["givens_real"][0] < 1e-14 and vals["givens_real"][1] == 0.0
    assert vals["givens_cplx"][0] < 1e-14 and vals["givens_cplx"][1] == 0.0 and vals["givens_cplx"][2] < 1e-14
