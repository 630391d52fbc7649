import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
REF_ROOT = "/root/reference"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: takes more than ~20 s on CPU")


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle
    pyoracle.build(ref=False)
    return pyoracle.Oracle()


@pytest.fixture(scope="session")
def ase_small():
    from raytrace_miniapp_b200 import problem_io
    return problem_io.load_npz(os.path.join(GOLDEN, "ase_small.npz"))


@pytest.fixture(scope="session")
def seed_small():
    from raytrace_miniapp_b200 import problem_io
    return problem_io.load_npz(os.path.join(GOLDEN, "seed_small.npz"))


@pytest.fixture(scope="session")
def rtlib():
    """The product library.  Built in-tree; GPU tests fail loudly when it is missing."""
    from raytrace_miniapp_b200 import build, lib
    build.build_library()
    return lib


@pytest.fixture(scope="session")
def ctx(rtlib):
    c = rtlib.Context(0)
    yield c
    c.close()


def rel_l2(a, b):
    a, b = np.asarray(a, float).ravel(), np.asarray(b, float).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def max_rel(a, b, floor=1e-6):
    """Max element-wise relative error over entries larger than floor*max."""
    a, b = np.asarray(a, float).ravel(), np.asarray(b, float).ravel()
    m = np.abs(b) > floor * np.abs(b).max()
    if not m.any():
        return 0.0
    return float(np.max(np.abs(a[m] - b[m]) / np.abs(b[m])))
