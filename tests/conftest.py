import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


@pytest.fixture(scope='session')
def cuda_platform():
    import torch
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    from atomsmm_b200 import mm
    return mm.Platform.getPlatformByName('B200')
