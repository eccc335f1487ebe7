"""Writes tests/golden/ref_krylov.npz (+ ref_solvers.npz): outputs of THE REFERENCE'S OWN FORTRAN SOURCES, executed statement by
statement by oracle/f90run.py (no Fortran compiler exists in this image or on the GPU box), on the inputs of
tests/golden/ref_cases.py.  Run in the container:   python tests/golden/make_ref_golden.py

The fixtures are what pins oracle/lk_oracle.c to the reference: tests/test_ref_golden.py compares the oracle (CPU) and the
product (GPU) with them.  /root/reference is needed to GENERATE them, never to run the tests.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)

import ref_cases as rc  # noqa: E402


def source_digest():
    """sha256 over the reference files the interpreter loaded -- recorded in the fixture so a reader can tell which text ran"""
    from oracle import ref_exec
    h = hashlib.sha256()
    for f in ref_exec.FILES:
        with open(os.path.join(ref_exec.REF, f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def generate(cases, path):
    be = rc.RefBackend()
    out = {}
    for name, fn in cases.items():
        for kind in rc.KINDS:
            if not rc.applies(name, kind):
                continue
            res = fn(kind, be)
            for key, val in res.items():
                out[f"{name}/{kind}/{key}"] = np.asarray(val)
            print(f"  {name:24s} {kind}  " + ", ".join(f"{k}={np.asarray(v).tolist()}" for k, v in res.items()
                                                        if np.asarray(v).ndim == 0))
    out["__reference_sha256__"] = np.array(source_digest())
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(out)} arrays, {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    generate(rc.CASES, os.path.join(HERE, "ref_krylov.npz"))
    if getattr(rc, "SOLVER_CASES", None):
        generate(rc.SOLVER_CASES, os.path.join(HERE, "ref_solvers.npz"))
