import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


@pytest.fixture(scope="session")
def golden():
    return load_golden


def split_cams(cams):
    """(C,12) camera rows -> (extrinsics (C,6), intrinsics list) in the reference's types."""
    intr = []
    for p in cams:
        K = np.eye(3)
        K[0, 0], K[1, 1], K[0, 2], K[1, 2] = p[:4]
        intr.append((K, np.r_[p[4:6], 0.0, 0.0, 0.0]))
    return cams[:, 6:].copy(), intr
