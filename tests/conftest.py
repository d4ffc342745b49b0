import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "ot-based-heterogeneous-multi-modal-fusion-embedding-for-ad-analysis-_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

for p in (os.path.join(ROOT, "tests"),):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def native_lib():
    """Path of libb200ot.so, building it with nvcc if the tree has no fresh copy.
    A missing library is a hard failure, never a skip: there is no fallback path."""
    from b200ot import build as b200ot_build
    return b200ot_build.build()


@pytest.fixture(scope="session")
def cuda_dev(native_lib):
    import torch
    assert torch.cuda.is_available(), "GPU tests were selected (-m gpu) but no CUDA device is visible"
    return torch.device("cuda", 0)


def pytest_sessionfinish(session, exitstatus):
    """Write the measured plan-parity figures (tests/_parity.py) where gpurun brings them back."""
    try:
        import json
        import _parity
        if _parity.REPORT:
            out = os.path.join(ROOT, "gpurun_out")
            os.makedirs(out, exist_ok=True)
            with open(os.path.join(out, "parity_report.json"), "w") as fh:
                json.dump(_parity.REPORT, fh, indent=1)
    except Exception:  # reporting must never turn a green run red
        pass
