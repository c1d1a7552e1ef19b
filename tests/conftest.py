import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_vectors.npz"), allow_pickle=False)


@pytest.fixture(scope="session")
def srs():
    """the product package, with libsrfuse.so built"""
    import __graft_entry__ as ge
    ge.build()
    import super_resolution_system_b200 as pkg
    return pkg


@pytest.fixture(scope="session")
def oracle():
    import oracle as o
    o.pyr.build()
    return o
