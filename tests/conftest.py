import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_PATH = os.path.join(ROOT, "tests", "golden", "reference_golden.npz")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden():
    """Outputs of the real reference on seeded inputs (oracle/make_golden.py)."""
    with np.load(GOLDEN_PATH) as z:
        return {k: z[k] for k in z.files}


def gcase(golden, name):
    pre = name + "/"
    return {k[len(pre):]: v for k, v in golden.items() if k.startswith(pre)}
