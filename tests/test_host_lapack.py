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
