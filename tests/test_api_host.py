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
