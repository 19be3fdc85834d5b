import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    path = os.path.join(ROOT, "tests", "golden", "reference_vectors.npz")
    return np.load(path)


@pytest.fixture(scope="session")
def sync_golden():
    """utils.synchronize_signals_improved of the unmodified reference (tests/golden/make_golden_sync.py)."""
    return np.load(os.path.join(ROOT, "tests", "golden", "sync_vectors.npz"))
