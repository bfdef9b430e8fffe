import os
import sys

import pytest

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (_REPO, os.path.join(_REPO, 'mrs-gym_b200')):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(_REPO, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


@pytest.fixture(scope='session')
def golden_dir():
    return GOLDEN
