import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def rel_l2(a, b):
    a = np.asarray(a); b = np.asarray(b)
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300))


@pytest.fixture(scope="session")
def protos():
    out = {}
    for M in (256, 512, 1024):
        p = np.load(os.path.join(GOLDEN, "prototype_M%d_m4_r1.npz" % M))
        out[M] = (p["h"], p["g"])
    return out


def load_golden(name):
    return np.load(os.path.join(GOLDEN, "golden_%s.npz" % name))
