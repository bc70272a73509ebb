import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box: pytest -m gpu)")


@pytest.fixture(scope="session")
def golden_tiny():
    return dict(np.load(os.path.join(GOLDEN, "pn2cls_tiny.npz")))


@pytest.fixture(scope="session")
def golden_full():
    return dict(np.load(os.path.join(GOLDEN, "pn2cls_full_2638.npz")))


@pytest.fixture(scope="session")
def cloud_2638():
    return np.load(os.path.join(GOLDEN, "cloud_2638_view0_25600.npy"))
