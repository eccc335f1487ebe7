"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/lkb.h declares, the ctypes table covers them all, and the product fails loudly (no CPU
fallback) when no CUDA device is present."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "lkb.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = set(re.findall(r"\b(lkb_[a-z0-9_]+)\s*\(", src))
    names -= {"lkb_matvec_fn"}
    return sorted(names)


@pytest.fixture(scope="module")
def so_path():
    from lightkrylov_b200 import build
    return build.build()


def test_header_symbols_are_exported(so_path):
    lib = ctypes.CDLL(so_path)
    syms = _declared_symbols()
    assert len(syms) >= 60
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, f"declared in include/lkb.h but not exported: {missing}"


def test_ctypes_table_matches_header(so_path):
    from lightkrylov_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared_symbols()
    _lib.load()


def test_no_cpu_fallback(so_path):
    """Without a GPU every entry point must fail loudly; nothing may silently run on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import lightkrylov_b200 as lk
    with pytest.raises(lk.LkbError, match="no CUDA device"):
        lk.Context(0)


def test_product_never_imports_oracle():
    """oracle/ is test infrastructure: the package and the CUDA sources must not reference it."""
    pkg = os.path.join(ROOT, "lightkrylov_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dp, f), errors="replace").read()
                assert "lk_oracle" not in txt and "liblk_oracle" not in txt and "from oracle" not in txt, f


def test_sass_has_128bit_loads(so_path):
    """The GS kernels must move data with 128-bit loads (LDG.E.128), checked on the built cubin."""
    import shutil, subprocess
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not available")
    out = subprocess.run(["cuobjdump", "-sass", so_path], capture_output=True, text=True).stdout
    assert "arch = sm_100a" in out
    # per-function check: the multi-dot and multi-axpy kernels stream with LDG.E.128
    for fn in ("k_multidot", "k_multiaxpy"):
        chunks = [c for c in out.split("Function : ")[1:] if fn in c.split("\n", 1)[0]]
        assert chunks, fn
        assert all("LDG.E.128" in c for c in chunks), fn


def _build_c_demo(so_path, tmp_path):
    import subprocess
    exe = str(tmp_path / "c_abi_demo")
    libdir = os.path.dirname(so_path)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "c_abi_demo.c"), "-L" + libdir, "-llkb",
                           "-Wl,-rpath," + libdir, "-lm", "-o", exe])
    return exe


def test_header_is_plain_c_and_links(so_path, tmp_path):
    """include/lkb.h must compile as C99 and a C client must link against liblkb.so (no C++/torch types)."""
    import subprocess
    exe = _build_c_demo(so_path, tmp_path)
    import torch
    if not torch.cuda.is_available():
        r = subprocess.run([exe, "64"], capture_output=True, text=True)
        assert r.returncode == 2 and "no CUDA device" in r.stderr      # loud failure, no CPU fallback


@pytest.mark.gpu
def test_c_client_runs_on_gpu(so_path, tmp_path):
    import subprocess
    exe = _build_c_demo(so_path, tmp_path)
    r = subprocess.run([exe, "1024"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "info=0" in r.stdout


def test_fortran_shim_binds_only_exported_symbols(so_path):
    """The ISO_C_BINDING shim cannot be compiled here (no Fortran compiler): at least every bind(C) name it
    uses must be an entry point of include/lkb.h that liblkb.so exports."""
    src = open(os.path.join(ROOT, "fortran", "lightkrylov_cuda.f90")).read()
    names = set(re.findall(r"bind\(C,\s*name='(lkb_[a-z0-9_]+)'\)", src))
    assert len(names) >= 20
    lib = ctypes.CDLL(so_path)
    declared = set(_declared_symbols())
    assert names <= declared, names - declared
    assert all(hasattr(lib, n) for n in names)
