"""GPU input pipeline (csrc/input_pipeline.cu + pixelssl_b200/task/sseg/gpu_input.py) against the fixture generated
by the UNMODIFIED reference's transform classes (tests/golden/input_pipeline.npz, oracle/make_golden.py:golden_input)
and against the oracle restatement of Pillow's arithmetic at the benchmark's crop size: bit for bit."""
import os
import random

import numpy as np
import pytest
import torch

from oracle import input_oracle as I

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), 'golden')


@pytest.fixture(scope='module')
def gin():
    if not torch.cuda.is_available():
        pytest.skip('needs a GPU')
    from pixelssl_b200.task.sseg import gpu_input
    return gpu_input


def test_train_prehandle_bit_exact_against_reference_golden(gin):
    g = np.load(os.path.join(G, 'input_pipeline.npz'))
    for k, (h, w, base, crop, labeled) in enumerate(g['cases']):
        random.seed(500 + k)
        x, y = gin.train_prehandle(g['img%d' % k], g['lab%d' % k] if labeled else None, int(base), int(crop))
        assert x.dtype == torch.float32 and tuple(x.shape) == (3, crop, crop)
        assert np.array_equal(x.cpu().numpy(), g['x%d' % k]), k
        assert np.array_equal(y.cpu().numpy(), g['y%d' % k]), k
        if not labeled:
            assert bool((y == -1.0).all())


@pytest.mark.parametrize('h,w,base,crop,seed', [(375, 500, 400, 513, 1), (500, 333, 400, 513, 2), (281, 500, 400, 321, 3),
                                                (120, 90, 400, 513, 4), (713, 713, 400, 713, 5)])
def test_train_prehandle_bit_exact_at_benchmark_sizes(gin, h, w, base, crop, seed):
    """Pascal-VOC-sized images, the reference scripts' base size 400 and the 513 / 713 crops: down- and up-scaling,
    padding (short edge < crop), both flip outcomes over the seeds."""
    rs = np.random.RandomState(seed)
    img = rs.randint(0, 256, (h, w, 3)).astype(np.uint8)
    lab = rs.randint(0, 21, (h, w)).astype(np.uint8)
    lab[rs.rand(h, w) < 0.1] = 255
    for labeled in (True, False):
        random.seed(100 + seed)
        x_ref, y_ref = I.train_prehandle(img, lab if labeled else None, base, crop)
        random.seed(100 + seed)
        x, y = gin.train_prehandle(img, lab if labeled else None, base, crop)
        assert np.array_equal(x.cpu().numpy(), x_ref)
        assert np.array_equal(y.cpu().numpy(), np.asarray(y_ref, dtype=np.float32))


@pytest.mark.parametrize('h,w,size,rescaling', [(37, 53, 33, True), (64, 41, 48, True), (30, 30, 30, True), (45, 70, 0, False),
                                                (375, 500, 513, True)])
def test_val_prehandle_bit_exact(gin, h, w, size, rescaling):
    rs = np.random.RandomState(h * 100 + w)
    img = rs.randint(0, 256, (h, w, 3)).astype(np.uint8)
    lab = rs.randint(0, 21, (h, w)).astype(np.uint8)
    x_ref, y_ref = I.val_prehandle(img, lab, size, rescaling)
    x, y = gin.val_prehandle(img, lab, size, rescaling)
    assert np.array_equal(x.cpu().numpy(), x_ref) and np.array_equal(y.cpu().numpy(), y_ref)
