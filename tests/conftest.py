import os
import sys

import numpy
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')
    # (re)build the in-tree C-ABI library and the oracle before any test module loads them
    from popdynamics_b200 import build
    if os.path.exists(build.NVCC):
        build.build_extension()
    from oracle import oracle
    oracle.build()


@pytest.fixture(scope='session')
def golden():
    path = os.path.join(ROOT, 'tests', 'golden', 'reference_vectors.npz')
    return numpy.load(path)


@pytest.fixture(scope='session')
def built_library():
    from popdynamics_b200 import build
    return build.build_extension()
