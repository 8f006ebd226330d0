import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def mnist():
    """The reference's own fixture mnist_test.csv.gz (tests/golden/mnist_test_u8.npz, made by
    scripts/make_mnist_fixture.py).  Returns (features float64 [10000,784], labels int32)."""
    z = np.load(os.path.join(ROOT, "tests", "golden", "mnist_test_u8.npz"))
    return z["pixels"].astype(np.float64), z["label"].astype(np.int32)
