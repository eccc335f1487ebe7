"""The reference's OWN unit tests (test/TestVectors.f90, TestLinops.f90, TestKrylov.f90, TestIterativeSolvers.f90,
TestExpmlib.f90), executed by oracle/f90run.py.  This is what qualifies the interpreter as a stand-in for the Fortran compiler
this image lacks: every `check(error, ...)` assertion the reference makes about its own arnoldi / lanczos / bidiagonalization /
qr / krylov_schur / eigs / eighs / svds / gmres / fgmres / cg / kexpm / expm / sqrtm must hold when its sources run here.
(test-drive's `check`, `new_unittest` and the reference's `check_test` reporting helper are the only things replaced.)

Needs /root/reference: skipped on the GPU box.  Default: the real(dp) instance of every test (about a minute);
LK_REF_SUITE=all runs all four kinds (4-5 minutes; last full run: profiles/r02_reference_suite_under_interpreter.txt).
"""
import os
import time

import numpy as np
import pytest

from oracle import f90run, ref_exec

TEST_FILES = ["TestVectors.f90", "TestLinops.f90", "TestKrylov.f90", "TestIterativeSolvers.f90", "TestExpmlib.f90"]

pytestmark = pytest.mark.skipif(not ref_exec.available(), reason="/root/reference not present on this box")


class _Checks:
    n = 0
    failed = 0


def _check(interp, error, cond=None, *rest, **kw):
    """test-drive check(error, condition) / check(error, actual, expected)"""
    _Checks.n += 1
    ok = bool(np.all(cond)) if not rest else bool(np.all(cond == rest[0]))
    if not ok:
        _Checks.failed += 1


@pytest.fixture(scope="module")
def suite():
    it = ref_exec.interp()
    for f in TEST_FILES:
        it.p.load(os.path.join(ref_exec.REF, "test", f))
    f90run.Interp(it.p)                      # evaluates the module-level parameters of the files just loaded
    it.natives.update({"check": _check, "check_test": f90run._n_noop, "get_err_str": f90run._n_noop,
                       "new_unittest": lambda interp, *a, **k: None})
    return it


def _tests(it, file_tag):
    names = [n for n, pr in it.p.procs.items() if n.startswith("test_") and pr.args == ["error"]
             and os.path.basename(pr.file) == file_tag]
    if os.environ.get("LK_REF_SUITE", "") != "all":
        names = [n for n in names if "_rdp" in n]
    return names


@pytest.mark.parametrize("file_tag", TEST_FILES)
def test_reference_tests_pass_under_the_interpreter(suite, file_tag):
    it = suite
    names = _tests(it, file_tag)
    assert names, f"no test procedures found in {file_tag}"
    report, bad = [], []
    for name in names:
        _Checks.n = _Checks.failed = 0
        it.rng = np.random.default_rng(7)
        t0 = time.time()
        try:
            it.call(name, None)
            status = "PASS" if _Checks.failed == 0 else f"ASSERTION FAILED ({_Checks.failed} of {_Checks.n})"
        except f90run.FortranError as exc:
            status = "ERROR " + str(exc).replace("\n", " | ")[:300]
        report.append(f"{name:48s} {status}  checks={_Checks.n}  {time.time() - t0:.1f}s")
        if status != "PASS":
            bad.append(report[-1])
    print("\n".join(report))
    assert not bad, "\n".join(bad)
