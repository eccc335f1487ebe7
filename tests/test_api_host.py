"""Host-side logic of the Python mirror that needs no GPU: kind mapping, in-place host matrix checks, tolerances."""
import numpy as np
import pytest


def test_kind_mapping_and_constants():
    import lightkrylov_b200 as lk
    assert [lk.kind_of(lk.DTYPES[k]) for k in "sdcz"] == list("sdcz")
    with pytest.raises(TypeError):
        lk.kind_of(np.int32)
    # Constants.f90:16-37 : atol = 10^-precision, rtol = sqrt(atol)
    assert lk.ATOL["d"] == 1e-15 and lk.ATOL["s"] == 1e-6
    assert abs(lk.RTOL["d"] - 3.1622776601683795e-08) < 1e-20 and abs(lk.RTOL["c"] - 1e-3) < 1e-12
    assert lk.KINDS == {"s": 0, "d": 1, "c": 2, "z": 3}


def test_host_matrix_must_be_fortran_ordered_and_typed():
    from lightkrylov_b200.api import _hostmat
    H = np.zeros((5, 4), order="F")
    assert _hostmat(H, "d") is H
    for bad in (np.zeros((5, 4), order="C"), np.zeros((5, 4), dtype=np.float32, order="F"), np.zeros(5), [[0.0]]):
        with pytest.raises(TypeError):
            _hostmat(bad, "d")


def test_signature_table_is_consistent_with_ctypes():
    import ctypes as C
    from lightkrylov_b200 import _lib
    for name, (res, args) in _lib.SIGNATURES.items():
        assert name.startswith("lkb_")
        assert res is None or isinstance(res, type) or res in (C.c_char_p,)
        assert isinstance(args, list)
    assert C.sizeof(_lib.GmresIO) == 48 and C.sizeof(_lib.CgIO) == 32      # must match lkb_gmres_io / lkb_cg_io in include/lkb.h


def test_write_results_format(tmp_path):
    """write_results (IterativeSolvers.fypp:882-924): header '(A6,4(A18),A6)', rows '(I6,4(2X,E16.9),2X,L4)', sorted by
    residual, first column = number of values.  Pure host code: runs without a GPU."""
    import lightkrylov_b200 as lk
    lam = np.array([1.5 + 2j, -0.25 + 0j, 1e-12 - 3e5j]); res = np.array([3e-7, 1e-9, 0.5])
    f = str(tmp_path / "eigs_output.txt")
    out = lk.write_results(f, lam, res, 1e-6)
    assert list(out) == [1e-9, 3e-7, 0.5]                                  # res comes back sorted, like the reference's
    lines = open(f).read().splitlines()
    assert lines[0] == "  Iter                Re                Im           modulus          residual  conv"
    assert lines[1] == "     3  -0.250000000E+00   0.000000000E+00   0.250000000E+00   0.100000000E-08     T"
    assert lines[2] == "     3   0.150000000E+01   0.200000000E+01   0.250000000E+01   0.300000000E-06     T"
    assert lines[3] == "     3   0.100000000E-11  -0.300000000E+06   0.300000000E+06   0.500000000E+00     F"
    f2 = str(tmp_path / "eighs_output.txt")
    lk.write_results(f2, np.array([11.99795014106112, 9.9999999996, 1.0]), res, 1e-6)
    lines = open(f2).read().splitlines()
    assert lines[0] == "  Iter             value          residual  conv"
    assert lines[1] == "     3   0.100000000E+02   0.100000000E-08     T"          # rounding carries into the exponent
    assert lines[2] == "     3   0.119979501E+02   0.300000000E-06     T"


def test_save_eigenspectrum_is_numpy_readable(tmp_path):
    """save_eigenspectrum (IterativeSolvers.fypp:941-960): .npy, Fortran order, (Re, Im, residual) or (value, residual)."""
    import lightkrylov_b200 as lk
    lam = np.array([1.5 + 2j, -0.25 + 0j, 1e-12 - 3e5j]); res = np.array([3e-7, 1e-9, 0.5])
    f = str(tmp_path / "spectrum.npy")
    lk.save_eigenspectrum(lam, res, f)
    a = np.load(f)
    assert a.dtype == np.float64 and a.shape == (3, 3) and a.flags["F_CONTIGUOUS"]
    assert np.array_equal(a[:, 0], lam.real) and np.array_equal(a[:, 1], lam.imag) and np.array_equal(a[:, 2], res)
    lk.save_eigenspectrum(lam.real.astype(np.float32), res, f)
    a = np.load(f)
    assert a.dtype == np.float32 and a.shape == (3, 2) and np.allclose(a[:, 1], res)
