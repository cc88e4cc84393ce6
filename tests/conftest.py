import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def pkg():
    from icp4r_loader import pkg as p
    return p


@pytest.fixture(scope="session")
def O():
    import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def handle(pkg):
    """One libicp4r_cuda handle on cuda:0. Fails (not skips) when the library or device is missing: the GPU
    tests must never pass on a fallback."""
    h = pkg.Icp4r(0)
    yield h
    h.close()


def rot_angle(R):
    """rotation angle of a 3x3 rotation matrix, accurate for tiny angles (arccos of the trace has a 1.5e-8 noise floor)"""
    import numpy as np
    R = np.asarray(R, np.float64)
    s = 0.5 * np.sqrt((R[2, 1] - R[1, 2]) ** 2 + (R[0, 2] - R[2, 0]) ** 2 + (R[1, 0] - R[0, 1]) ** 2)
    return float(np.arctan2(s, (np.trace(R) - 1) / 2))
