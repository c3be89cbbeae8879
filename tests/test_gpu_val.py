"""GPU parity of the validation-side kernels: confusion matrix (bit-exact integers), the segmentation
metrics through the TaskFunc hook, and the Mean-Teacher Gaussian input-noise layer (bit-exact fp32 for
a given noise tensor).  Goldens come from the unmodified reference (oracle/make_golden.py val)."""
import os
import types

import numpy as np
import pytest
import torch

from oracle import sseg_oracle as O

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), 'golden')


@pytest.fixture(scope='module')
def ops():
    if not torch.cuda.is_available():
        pytest.skip('needs a GPU')
    from pixelssl_b200 import ops as _ops
    return _ops


def test_confusion_matrix_bit_exact_vs_reference(ops):
    g = np.load(os.path.join(G, 'val.npz'))
    cmat = torch.zeros(21, 21, dtype=torch.int64, device='cuda')
    for k in range(2):
        ops.confusion_matrix_(cmat, torch.from_numpy(g['metrics_pred%d' % k]).cuda(),
                              torch.from_numpy(g['metrics_gt%d' % k]).cuda(), 21)
        assert np.array_equal(cmat.cpu().numpy(), g['metrics_cmat_sum%d' % k])


def test_metrics_hook_matches_reference_values(ops):
    from pixelssl_b200.task.sseg.func import SemanticSegmentationFunc
    from pixelssl_b200.utils import logger
    g = np.load(os.path.join(G, 'val.npz'))
    tf = SemanticSegmentationFunc(types.SimpleNamespace(num_classes=21))
    meters = logger.AvgMeterSet()
    for k in range(2):
        tf.metrics((torch.from_numpy(g['metrics_pred%d' % k]).cuda(),),
                   (torch.from_numpy(g['metrics_gt%d' % k]).cuda(),), None, meters, id_str='task')
        got = np.array([float(meters['task_metric_' + m].val) for m in ('acc', 'acc-class', 'mIoU', 'fwIoU')])
        np.testing.assert_allclose(got, g['metrics_values%d' % k], rtol=1e-12)
        assert np.array_equal(np.array(meters['task_confusion_matrix'].sum), g['metrics_cmat_sum%d' % k])


@pytest.mark.parametrize('n,c,h,w', [(1, 2, 1, 1), (2, 21, 7, 300), (4, 21, 513, 513), (3, 64, 33, 35)])
def test_confusion_matrix_vs_oracle_and_properties(ops, n, c, h, w):
    gs = torch.Generator().manual_seed(n * 1000 + h)
    pred = torch.randn(n, c, h, w, generator=gs)
    pred[:, :, ::5, ::7] = 0.0                                  # ties everywhere on a sub-grid -> class 0
    gt = torch.randint(0, c, (n, 1, h, w), generator=gs).float()
    gt[torch.rand(n, 1, h, w, generator=gs) < 0.07] = 255.0
    gt[torch.rand(n, 1, h, w, generator=gs) < 0.01] = -1.0
    want = O.confusion_matrix(pred.numpy(), gt.numpy(), c)
    cmat = torch.zeros(c, c, dtype=torch.int64, device='cuda')
    ops.confusion_matrix_(cmat, pred.cuda(), gt.cuda())
    got = cmat.cpu().numpy()
    assert np.array_equal(got, want)
    valid = int(((gt >= 0) & (gt < c)).sum())
    assert got.sum() == valid                                   # every valid pixel counted exactly once
    assert np.array_equal(got.sum(axis=1), np.bincount(gt[(gt >= 0) & (gt < c)].long().numpy(), minlength=c))
    ops.confusion_matrix_(cmat, pred.cuda(), gt.cuda())         # accumulation is linear
    assert np.array_equal(cmat.cpu().numpy(), 2 * want)


def test_confusion_matrix_rejects_bad_arguments(ops):
    pred = torch.zeros(1, 3, 4, 4, device='cuda')
    gt = torch.zeros(1, 1, 4, 4, device='cuda')
    with pytest.raises(TypeError):
        ops.confusion_matrix_(torch.zeros(3, 3, device='cuda'), pred, gt)
    with pytest.raises(ValueError):
        ops.confusion_matrix_(torch.zeros(3, 3, dtype=torch.int64, device='cuda'), pred, gt[:, :, :2])


def test_gaussian_noise_bit_exact_vs_reference(ops):
    g = np.load(os.path.join(G, 'val.npz'))
    x = torch.from_numpy(g['gn_inp']).cuda()
    out = ops.gaussian_noise_(x, 0.15, noise=torch.from_numpy(g['gn_noise']).cuda())
    assert out.data_ptr() == x.data_ptr()                       # in place like the reference layer
    assert np.array_equal(out.cpu().numpy(), g['gn_out'])


@pytest.mark.parametrize('shape', [(1, 1, 1, 1), (2, 3, 17, 5), (16, 3, 513, 513)])
def test_gaussian_noise_vs_oracle_and_range(ops, shape):
    gs = torch.Generator().manual_seed(shape[-1])
    x = torch.randn(shape, generator=gs) * 3.0 + 1.0
    noise = torch.randn(shape, generator=gs) * 0.3
    want = O.gaussian_noise_layer(x, noise)
    got = ops.gaussian_noise_(x.clone().cuda(), 0.3, noise=noise.cuda()).cpu()
    assert torch.equal(got, want)
    # clipping keeps every sample inside its own [min, max]
    lo = x.amin(dim=(1, 2, 3), keepdim=True)
    hi = x.amax(dim=(1, 2, 3), keepdim=True)
    assert bool(((got >= lo - 1e-5) & (got <= hi + 1e-5)).all())
    # zero noise is (nearly) the identity, disabled layer is exactly the identity
    ident = ops.gaussian_noise_(x.clone().cuda(), 0.3, noise=torch.zeros(shape).cuda()).cpu()
    assert float((ident - x).abs().max()) <= 1e-5 * float(x.abs().max()) + 1e-9
    same = x.clone().cuda()
    assert ops.gaussian_noise_(same, None) is same and torch.equal(same.cpu(), x)


def test_gaussian_noise_draws_independent_noise(ops):
    torch.manual_seed(0)
    x = torch.rand(2, 3, 33, 33, device='cuda')
    a = ops.gaussian_noise_(x.clone(), 0.2)
    b = ops.gaussian_noise_(x.clone(), 0.2)
    assert not torch.equal(a, b)
    assert float((a - x).abs().max()) < 2.0


def _cfg(alg, **kw):
    cfg = {'ssl_algorithm': alg, 'lr': 0.00025, 'momentum': 0.9, 'weight_decay': 0.0005,
           'output_stride': 16, 'backbone': 'resnet101', 'epochs': 2, 'log_freq': 1000}
    cfg.update(kw)
    return cfg


def test_mt_train_with_input_noise_and_validate_metrics(ops):
    """gaussian_noise_std: student and teacher read differently-noised copies (ssl_mt.py:337-349);
    ``_validate`` fills the reference's metric meters (ssl_mt.py:264-265, func.py:50-80) and its
    confusion matrix equals the oracle's on the model's own predictions."""
    from pixelssl_b200 import runner
    ops.set_conv_precision('fp32')
    args = runner.build_args(_cfg('ssl_mt', cons_for_labeled=False, cons_scale=1.0, cons_rampup_epochs=1,
                                  ema_decay=0.99, batch_size=4, unlabeled_batch_size=2, gaussian_noise_std=0.15),
                             iters_per_epoch=5)
    alg = runner.build_algorithm(args)
    img, lab = O.synthetic_batch(3, 4, 2, 65, 65)
    s_inp, t_inp, gt = alg._batch_prehandle((img,), (lab,), True)
    assert not torch.equal(s_inp[0], t_inp[0]) and not torch.equal(s_inp[0].cpu(), img)
    assert float((s_inp[0].cpu() - img).abs().max()) < 0.15 * 6 * float(img.max() - img.min())
    v_inp, v_t_inp, _ = alg._batch_prehandle((img,), (lab,), False)
    assert v_inp[0] is v_t_inp[0] and torch.equal(v_inp[0].cpu(), img)
    alg._train([((img,), (lab,))], 0)
    assert np.isfinite(float(alg.meters['s_task_loss'].val)) and np.isfinite(float(alg.meters['cons_loss'].val))

    lab_full = lab.clone()
    lab_full[2:] = lab[:2]                                       # validation batches are fully labeled
    alg._validate([((img,), (lab_full,))], 0)
    for who in ('student', 'teacher'):
        for m in ('acc', 'acc-class', 'mIoU', 'fwIoU'):
            v = float(alg.meters['%s_metric_%s' % (who, m)].val)
            assert 0.0 <= v <= 1.0
    alg.s_model.eval()
    with torch.no_grad():
        res, _ = alg.s_model.forward((img.cuda(),))
    want = O.confusion_matrix(res['activated_pred'][0].cpu().numpy(), lab_full.numpy(), args.num_classes)
    assert np.array_equal(np.array(alg.meters['student_confusion_matrix'].sum), want)
