import os

# before anything creates the CUDA context (mogp_b200/__init__.py explains): one hardware queue per stream of the pipeline
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
if os.path.join(ROOT, "tests") not in sys.path:
    sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on a B200 through gpurun)")
    config.addinivalue_line("markers", "slow: longer CPU test")


@pytest.fixture(scope="session")
def single_blas_thread():
    """Small-matrix oracle runs are faster (and deterministic) with one BLAS thread."""
    from threadpoolctl import threadpool_limits
    with threadpool_limits(1):
        yield


def tutorial_data():
    """The reference tutorial's toy data + onset anchor
    (example/tutorial_train_mogp_model.ipynb:L52-58, mogp/utils.py:32-47)."""
    import numpy as np
    import oracle
    X, Y = oracle.generate_toy_data(seed=0)
    n = X.shape[0]
    X = np.hstack((np.zeros((n, 1)), X))
    Y = np.hstack((48 * np.ones((n, 1)), Y))
    return X, Y


# Reference-equivalent k-means initial partition of the tutorial: sklearn 1.9
# returns the same partition with labels 0<->1 swapped, so the KATs inject the reference numbering.
TUTORIAL_Z0_BLOCK = [1, 2, 2, 0, 0, 3, 2, 1, 2, 3, 0, 0, 4, 0, 4, 0, 0, 0, 0, 0]
