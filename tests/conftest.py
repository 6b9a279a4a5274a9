import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")

# fp32 parity contract from BASELINE.json `north_star`: 1e-5 relative / 1e-6 absolute
RTOL = 1e-5
ATOL = 1e-6


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    def load(name):
        return dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    return load


def pytest_sessionfinish(session, exitstatus):
    """Persist the element-wise parity report (tests/parity.py) of a GPU run: gpurun_out/parity_report.txt travels back from
    the GPU box and is copied to profiles/ by hand."""
    try:
        import parity
    except ImportError:
        return
    if parity.REPORT:
        out = os.path.join(ROOT, "gpurun_out")
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity_report.txt"), "w") as f:
            f.write("\n".join(parity.REPORT) + "\n")
