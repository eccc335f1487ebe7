import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _has_cuda() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """On a box without a CUDA device the GPU tests are SKIPPED (not failed / errored), so a CPU-only CI run of the
    whole directory distinguishes regressions from the missing device.  The explicit "no CPU fallback, fails loudly"
    check lives in tests/test_abi.py::test_no_cpu_fallback and runs exactly there."""
    if _has_cuda():
        return
    skip = pytest.mark.skip(reason="no CUDA device on this box (GPU parity tests run with -m gpu on a B200)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from oracle import lk_oracle
    lk_oracle.lib()
    return lk_oracle
