import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests', 'hostemu')):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200)")


GOLD = os.path.join(ROOT, 'tests', 'golden')


@pytest.fixture(scope='session')
def gold():
    import numpy as np

    def load(name):
        return np.load(os.path.join(GOLD, name + '.npz'))
    return load
