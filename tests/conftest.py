import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')

# Convolution precision modes every whole-network / whole-step GPU golden is run under (fixture params, so the
# mode shows up in the test id): the exact-fp32 FFMA kernels and the fp32-grade tcgen05 modes that bench.py times.
TEST_PRECISIONS = [p for p in os.environ.get('PXL_TEST_PRECISIONS', 'fp32,tf32x3').split(',') if p]


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')
    config.addinivalue_line('markers', 'slow: CPU test that takes more than a few seconds')


@pytest.fixture(scope='session')
def golden_dir():
    return GOLDEN
