"""CPU oracle for the PixelSSL sseg SSL-training hot path.  TEST INFRASTRUCTURE ONLY.

This file is the *checker*, never the product: only ``tests/``, ``__graft_entry__.smoke()``
and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  Nothing
under ``pixelssl_b200/`` imports it, and the product path raises if its CUDA library is absent.

What it is: a functional fp32 restatement (torch CPU ops + numpy, no nn.Module graph) of the
reference's per-step algorithm.  The reference's arithmetic lives in an un-vendored dependency
(PyTorch, unpinned: ``pixelssl/requirements.txt`` lists no torch; ``runner.py:29`` asks >= 1.0.0);
the de-facto pin is the torch 2.11 CPU fp32 kernels in this image, which is what this file calls.

Parity pin: ``oracle/make_golden.py`` imports the real reference from ``/root/reference`` in the
build container, drives the unmodified ``SSLMT._train`` / ``SSLNULL._train`` /
``SSLCUTMIX._train`` bodies and op call sites on seeded inputs and commits the results under
``tests/golden/``; ``tests/test_oracle_golden.py`` holds this restatement to those vectors.  The
reference ships no tests / golden vectors of its own, so reference-generated
fixtures are the pin.

Every function cites the reference file:line it restates (paths relative to the reference root).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------------------------
# network description (task/sseg/module/backbone/resnet.py:52-131, deeplab_v2.py:13-85)
# ----------------------------------------------------------------------------------------------

R101_BLOCKS = (3, 4, 23, 3)


def resnet_plan(output_stride=16, blocks=R101_BLOCKS):
    """Yield (prefix, inplanes, planes, stride, dilation, has_downsample) per bottleneck.

    resnet.py:58-66 (stride / dilation tables), :85-119 (_make_layer, _make_MG_unit: layer4 uses
    the multi-grid unit [1,2,4] x dilation and always exactly 3 blocks)."""
    if output_stride == 16:
        strides, dilations = (1, 2, 2, 1), (1, 1, 1, 2)
    elif output_stride == 8:
        strides, dilations = (1, 2, 1, 1), (1, 1, 2, 4)
    else:
        raise NotImplementedError
    plan = []
    inplanes = 64
    for li, planes in enumerate((64, 128, 256, 512)):
        if li < 3:
            dil = [dilations[li]] * blocks[li]
        else:
            dil = [m * dilations[li] for m in (1, 2, 4)]
        for bi, d in enumerate(dil):
            stride = strides[li] if bi == 0 else 1
            down = bi == 0 and (stride != 1 or inplanes != planes * 4)
            plan.append(('layer%d.%d' % (li + 1, bi), inplanes, planes, stride, d, down))
            inplanes = planes * 4
    return plan


def deeplabv2_param_shapes(num_classes=21, output_stride=16, blocks=R101_BLOCKS):
    """Ordered (name, shape, kind) list in ``nn.Module.parameters()`` / ``state_dict`` order.

    kind in {'conv', 'bn_w', 'bn_b', 'cls_w', 'cls_b'}; buffers (running stats) are listed by
    :func:`deeplabv2_buffer_shapes`.  Names are relative to ``DeepLabV2`` (deeplab_v2.py:13), i.e.
    the checkpoint key is ``module.model.<name>`` (task/sseg/model.py:40, nn/func.py:58)."""
    out = [('backbone.conv1.weight', (64, 3, 7, 7), 'conv'),
           ('backbone.bn1.weight', (64,), 'bn_w'), ('backbone.bn1.bias', (64,), 'bn_b')]
    for prefix, inpl, pl, stride, dil, down in resnet_plan(output_stride, blocks):
        p = 'backbone.' + prefix
        out += [(p + '.conv1.weight', (pl, inpl, 1, 1), 'conv'),
                (p + '.bn1.weight', (pl,), 'bn_w'), (p + '.bn1.bias', (pl,), 'bn_b'),
                (p + '.conv2.weight', (pl, pl, 3, 3), 'conv'),
                (p + '.bn2.weight', (pl,), 'bn_w'), (p + '.bn2.bias', (pl,), 'bn_b'),
                (p + '.conv3.weight', (pl * 4, pl, 1, 1), 'conv'),
                (p + '.bn3.weight', (pl * 4,), 'bn_w'), (p + '.bn3.bias', (pl * 4,), 'bn_b')]
        if down:
            out += [(p + '.downsample.0.weight', (pl * 4, inpl, 1, 1), 'conv'),
                    (p + '.downsample.1.weight', (pl * 4,), 'bn_w'),
                    (p + '.downsample.1.bias', (pl * 4,), 'bn_b')]
    for i in range(4):
        out += [('classifier.conv2d_list.%d.weight' % i, (num_classes, 2048, 3, 3), 'cls_w'),
                ('classifier.conv2d_list.%d.bias' % i, (num_classes,), 'cls_b')]
    return out


def deeplabv2_buffer_shapes(output_stride=16, blocks=R101_BLOCKS):
    """BN running_mean / running_var / num_batches_tracked names, one triple per BN layer."""
    names = ['backbone.bn1']
    for prefix, inpl, pl, stride, dil, down in resnet_plan(output_stride, blocks):
        p = 'backbone.' + prefix
        names += [p + '.bn1', p + '.bn2', p + '.bn3']
        if down:
            names.append(p + '.downsample.1')
    chans = {}
    for n, shp, kind in deeplabv2_param_shapes(21, output_stride, blocks):
        if kind == 'bn_w':
            chans[n[:-len('.weight')]] = shp[0]
    return [(n, chans[n]) for n in names]


def init_deeplabv2(seed, num_classes=21, output_stride=16, blocks=R101_BLOCKS, cls_bias_std=0.0):
    """Deterministic state (params + BN buffers) with the reference's initial *distributions*.

    resnet.py:133-143: conv ~ N(0, sqrt(2/(kh*kw*out))), BN gamma=1 beta=0;
    deeplab_v2.py:78-79: classifier weight ~ N(0, 0.01) (bias keeps nn.Conv2d's default uniform
    init in the reference; here it is drawn N(0, cls_bias_std), 0 by default - the goldens load
    this very state into the reference model, so the draw itself need not match nn.Conv2d's).
    Draw order is fixed (parameter order, one torch.Generator), so the same seed gives the same
    weights in the golden generator, the CPU tests and on the GPU box."""
    g = torch.Generator().manual_seed(seed)
    state = {}
    for name, shape, kind in deeplabv2_param_shapes(num_classes, output_stride, blocks):
        if kind == 'conv':
            std = math.sqrt(2.0 / (shape[2] * shape[3] * shape[0]))
            state[name] = torch.randn(shape, generator=g) * std
        elif kind == 'cls_w':
            state[name] = torch.randn(shape, generator=g) * 0.01
        elif kind == 'cls_b':
            state[name] = torch.randn(shape, generator=g) * cls_bias_std
        elif kind == 'bn_w':
            state[name] = torch.ones(shape)
        else:
            state[name] = torch.zeros(shape)
    for name, c in deeplabv2_buffer_shapes(output_stride, blocks):
        state[name + '.running_mean'] = torch.zeros(c)
        state[name + '.running_var'] = torch.ones(c)
        state[name + '.num_batches_tracked'] = torch.tensor(0, dtype=torch.long)
    return state


def randomize_bn_affine(state, seed):
    """Perturb BN gamma/beta away from (1, 0) so backward parity tests exercise them."""
    g = torch.Generator().manual_seed(seed)
    for k in list(state.keys()):
        if k.endswith('.weight') and state[k].dim() == 1:
            state[k] = 1.0 + 0.2 * torch.randn(state[k].shape, generator=g)
        elif k.endswith('.bias') and state[k].dim() == 1 and 'classifier' not in k:
            state[k] = 0.1 * torch.randn(state[k].shape, generator=g)
    return state


# ----------------------------------------------------------------------------------------------
# forward pieces
# ----------------------------------------------------------------------------------------------

def batch_norm(x, state, prefix, training, momentum=0.1, eps=1e-5):
    """sync_batchnorm/batchnorm.py:48-53: the single-replica / eval path is ``F.batch_norm``
    (biased var + eps for normalisation, unbiased var into running_var, momentum 0.1, eps 1e-5
    from :39).  Mutates the running buffers in ``state`` in place, like the module does."""
    rm, rv = state[prefix + '.running_mean'], state[prefix + '.running_var']
    y = F.batch_norm(x, rm, rv, state[prefix + '.weight'], state[prefix + '.bias'],
                     training, momentum, eps)
    if training and (prefix + '.num_batches_tracked') in state:
        state[prefix + '.num_batches_tracked'] += 1
    return y


def sync_batch_norm_multi_replica(x_parts, weight, bias, running_mean, running_var,
                                  momentum=0.1, eps=1e-5):
    """sync_batchnorm/batchnorm.py:55-78,113-125: the >1-replica training path.  Per replica
    sum / square-sum, reduced; ``inv_std = clamp(biased_var, eps) ** -0.5`` (NOT var+eps) and
    running_var takes the unbiased variance.  Returns the per-replica outputs and the new
    running stats.  Used to pin the N>1 semantics of the engine's NCCL-synced BN."""
    c = weight.numel()
    size = sum(p.numel() // c for p in x_parts)
    s = sum(p.transpose(0, 1).reshape(c, -1).sum(1) for p in x_parts)
    ss = sum((p ** 2).transpose(0, 1).reshape(c, -1).sum(1) for p in x_parts)
    mean = s / size
    sumvar = ss - s * mean
    unbias_var, bias_var = sumvar / (size - 1), sumvar / size
    new_rm = (1 - momentum) * running_mean + momentum * mean
    new_rv = (1 - momentum) * running_var + momentum * unbias_var
    inv_std = bias_var.clamp(eps) ** -0.5
    outs = [(p - mean.view(1, c, 1, 1)) * (inv_std * weight).view(1, c, 1, 1) + bias.view(1, c, 1, 1)
            for p in x_parts]
    return outs, new_rm, new_rv


def bottleneck(x, state, p, stride, dilation, down, training):
    """resnet.py:30-50: 1x1 -> BN -> ReLU -> 3x3(stride, dilation, pad=dilation) -> BN -> ReLU ->
    1x1 -> BN -> (+ residual or downsample(x)) -> ReLU; every conv bias-free (:18-25)."""
    out = F.conv2d(x, state[p + '.conv1.weight'])
    out = F.relu(batch_norm(out, state, p + '.bn1', training))
    out = F.conv2d(out, state[p + '.conv2.weight'], stride=stride, padding=dilation, dilation=dilation)
    out = F.relu(batch_norm(out, state, p + '.bn2', training))
    out = F.conv2d(out, state[p + '.conv3.weight'])
    out = batch_norm(out, state, p + '.bn3', training)
    if down:
        res = F.conv2d(x, state[p + '.downsample.0.weight'], stride=stride)
        res = batch_norm(res, state, p + '.downsample.1', training)
    else:
        res = x
    return F.relu(out + res)


def resnet_forward(x, state, training, output_stride=16, blocks=R101_BLOCKS, prefix='backbone.'):
    """resnet.py:121-131: conv1 7x7/2 pad 3 -> BN -> ReLU -> maxpool 3x3/2 pad 1 -> layer1..4."""
    x = F.conv2d(x, state[prefix + 'conv1.weight'], stride=2, padding=3)
    x = F.relu(batch_norm(x, state, prefix + 'bn1', training))
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    for bp, inpl, pl, stride, dil, down in resnet_plan(output_stride, blocks):
        x = bottleneck(x, state, prefix + bp, stride, dil, down, training)
    return x


def aspp_classifier(x, state, prefix='classifier.conv2d_list.', dilations=(6, 12, 18, 24)):
    """deeplab_v2.py:71-85: sum over d of conv3x3(2048 -> C, dilation=d, padding=d, bias=True)."""
    out = None
    for i, d in enumerate(dilations):
        y = F.conv2d(x, state['%s%d.weight' % (prefix, i)], state['%s%d.bias' % (prefix, i)],
                     padding=d, dilation=d)
        out = y if out is None else out + y
    return out


def bilinear_align_corners(x, size):
    """deeplab_v2.py:32: ``F.interpolate(x, size, mode='bilinear', align_corners=True)``."""
    return F.interpolate(x, size=size, mode='bilinear', align_corners=True)


def deeplabv2_forward(img, state, training=True, output_stride=16, blocks=R101_BLOCKS):
    """deeplab_v2.py:29-33 + task/sseg/model.py:50-65.  Returns (logits, latent)."""
    latent = resnet_forward(img, state, training, output_stride, blocks)
    low = aspp_classifier(latent, state)
    return bilinear_align_corners(low, img.shape[2:]), latent


def channel_softmax(logits):
    """task/sseg/model.py:62: ``F.softmax(pred, dim=1)`` (the 'activated_pred')."""
    return F.softmax(logits, dim=1)


# ----------------------------------------------------------------------------------------------
# losses
# ----------------------------------------------------------------------------------------------

def sseg_criterion(logits, gt, ignore_index=255):
    """task/sseg/criterion.py:24-38: per-pixel CE (ignore_index, reduction none) on
    ``gt.long()`` then mean over ALL H*W pixels per sample (ignored pixels contribute 0 to the
    numerator but still count in the denominator) -> Tensor[n]."""
    n, c, h, w = logits.shape
    if gt.dim() == 4:
        gt = gt.view(n, h, w)
    loss = F.cross_entropy(logits, gt.long(), ignore_index=ignore_index, reduction='none')
    return loss.mean(dim=(1, 2))


def mse_consistency(student, teacher):
    """ssl_mt.py:115,179-187: ``nn.MSELoss()`` (mean over every element); teacher is detached."""
    return F.mse_loss(student, teacher.detach())


def sigmoid_rampup(current, rampup_length):
    """nn/func.py:12-20."""
    if rampup_length == 0:
        return 1.0
    current = float(np.clip(current, 0.0, rampup_length))
    phase = 1.0 - current / rampup_length
    return float(np.exp(-5.0 * phase * phase))


def bce_with_logits_mean(pred, target):
    """ssl_adv.py:496-503 FCDiscriminatorCriterion: BCE-with-logits, mean over (1,2,3) -> [B]."""
    return F.binary_cross_entropy_with_logits(pred, target, reduction='none').mean(dim=(1, 2, 3))


# ----------------------------------------------------------------------------------------------
# optimiser / schedule / EMA
# ----------------------------------------------------------------------------------------------

def poly_lr(base_lr, cur_iter, max_iters, power=0.9):
    """nn/lrer.py:156-158: ``base * (1 - cur_iter / max_iters) ** power``.  NOTE (:173-176 +
    torch's _LRScheduler.__init__ calling step() once): the first optimiser step already runs
    with cur_iter == 1."""
    return base_lr * ((1 - float(cur_iter) / max_iters) ** power)


def sgd_momentum_step(params, grads, bufs, lrs, momentum, weight_decay, first_step):
    """torch.optim.SGD as configured by nn/optimizer.py:57-75 (dampening 0, nesterov False):
    g += wd * p ; buf = g (first step) else mom * buf + g ; p -= lr * buf.  In place."""
    for p, g, b, lr in zip(params, grads, bufs, lrs):
        d = g + weight_decay * p if weight_decay != 0 else g.clone()
        if momentum != 0:
            if first_step:
                b.copy_(d)
            else:
                b.mul_(momentum).add_(d)
            d = b
        p.add_(d, alpha=-lr)


def ema_update(t_params, s_params, ema_decay, cur_step):
    """ssl_mt.py:359-363 / ssl_cutmix.py:434-438: d = min(1 - 1/(step+1), ema_decay);
    t = t*d + (1-d)*s over ``parameters()`` only (BN buffers are NOT averaged)."""
    d = min(1 - 1 / (cur_step + 1), ema_decay)
    for t, s in zip(t_params, s_params):
        t.mul_(d).add_(s, alpha=1 - d)
    return d


def lr_multipliers(names):
    """task/sseg/model.py:45-48 + deeplab_v2.py:42-60: backbone 1x, classifier 10x."""
    return [10.0 if n.startswith('classifier.') else 1.0 for n in names]


# ----------------------------------------------------------------------------------------------
# CutMix host pieces (bit-exact parts of the path)
# ----------------------------------------------------------------------------------------------

def box_masks(rng, mask_num, mask_shape, prop_range=(0.5, 0.5), boxes_num=1, invert=True):
    """ssl_cutmix.py:470-547 with the configuration the algorithm constructs (:126-128):
    area_prop, random aspect ratio, within bounds, invert -> 1 inside the box.  ``rng`` is a
    ``numpy.random.RandomState`` (the reference uses the global ``np.random``; same stream for
    the same seed).  Draw order: proportions, aspect exponents, positions."""
    props = rng.uniform(prop_range[0], prop_range[1], size=(mask_num, boxes_num))
    zero = props == 0.0
    y_props = np.exp(rng.uniform(low=0.0, high=1.0, size=(mask_num, boxes_num)) * np.log(props))
    x_props = props / y_props
    fac = np.sqrt(1.0 / boxes_num)
    y_props *= fac
    x_props *= fac
    y_props[zero] = 0
    x_props[zero] = 0
    shape = np.array(mask_shape)
    sizes = np.round(np.stack([y_props, x_props], axis=2) * shape[None, None, :])
    positions = np.round((shape - sizes) * rng.uniform(low=0.0, high=1.0, size=sizes.shape))
    rects = np.append(positions, positions + sizes, axis=2)
    masks = np.zeros((mask_num, 1) + tuple(mask_shape)) if invert else np.ones((mask_num, 1) + tuple(mask_shape))
    for i, sample in enumerate(rects):
        for y0, x0, y1, x1 in sample:
            sl = (i, 0, slice(int(y0), int(y1)), slice(int(x0), int(x1)))
            masks[sl] = 1 - masks[sl]
    return masks.astype(np.float32), rects


def cutmix_mix(mask, a, b):
    """ssl_cutmix.py:195,428: ``mask * a + (1 - mask) * b`` in fp32, each op rounded separately
    (no FMA contraction) - the engine kernel must be bit-identical to this."""
    return mask * a + (1 - mask) * b


def cutmix_confidence(mixed_prob, threshold):
    """ssl_cutmix.py:200: one scalar for the whole batch: mean(max_c p > thr)."""
    return (mixed_prob.max(dim=1)[0] > threshold).float().mean()


# ----------------------------------------------------------------------------------------------
# GaussianNoise (MT input perturbation) and Gaussian blur (GCT)
# ----------------------------------------------------------------------------------------------

def gaussian_noise(inp, noise):
    """nn/module/gaussian_noise.py:25-40 with the noise tensor injected (the reference draws
    ``normal_(0, random.uniform(0, std))``, :23).  Returns a new tensor (reference is in place)."""
    x = inp.clone()
    imax = x.amax(dim=(1, 2, 3), keepdim=True)
    imin = x.amin(dim=(1, 2, 3), keepdim=True)
    x.sub_(imin).div_(imax - imin + 1e-9)
    x.add_(noise)
    ub = (x > 1.0).float()
    lb = (x < 0.0).float()
    x.mul_(1 - ub).add_(ub)
    x.mul_(1 - lb)
    x.mul_(imax - imin + 1e-9).add_(imin)
    return x


def gaussian_kernel_1d(k):
    """nn/module/gaussian_blur.py:52-64: the reference filters a k x k delta with
    ``scipy.ndimage.gaussian_filter(sigma=0.3*((k-1)*0.5-1)+0.8)``; that 2-D kernel is exactly
    outer(v, v) with v the 1-D filtered delta (scipy truncates at 4 sigma, 'reflect' mode)."""
    from scipy.ndimage import gaussian_filter1d
    sigma = 0.3 * ((k - 1) * 0.5 - 1) + 0.8
    d = np.zeros(k)
    d[k // 2] = 1
    return gaussian_filter1d(d, sigma)


def gaussian_kernel_2d(k):
    """gaussian_blur.py:52-64 restated literally (2-D filter of a delta), fp32 like the layer."""
    from scipy.ndimage import gaussian_filter
    sigma = 0.3 * ((k - 1) * 0.5 - 1) + 0.8
    n = np.zeros((k, k))
    n[k // 2, k // 2] = 1
    return gaussian_filter(n, sigma).astype(np.float32)


def gaussian_blur(x, k):
    """gaussian_blur.py:30-50: ReflectionPad2d(k//2) then depthwise k x k conv, per channel."""
    c = x.shape[1]
    w = torch.from_numpy(gaussian_kernel_2d(k))[None, None].repeat(c, 1, 1, 1)
    return F.conv2d(F.pad(x, (k // 2,) * 4, mode='reflect'), w, groups=c)


# ----------------------------------------------------------------------------------------------
# whole steps
# ----------------------------------------------------------------------------------------------

class MTOracle:
    """Functional restatement of ``SSLMT._train``'s loop body (ssl_mt.py:131-220) plus the
    supervised-only variant (ssl_null.py:78-144) for DeepLabV2 on CPU.

    Holds student / teacher states (dict name -> tensor) in reference parameter order, SGD
    momentum buffers, and the PolynomialLR iteration counter."""

    def __init__(self, s_state, t_state=None, lr=2.5e-4, momentum=0.9, weight_decay=5e-4,
                 max_iters=1000, power=0.9, cons_scale=1.0, rampup_steps=0, ema_decay=0.99,
                 cons_for_labeled=False, num_classes=21, output_stride=16, blocks=R101_BLOCKS,
                 ignore_index=255):
        self.s, self.t = s_state, t_state
        self.names = [n for n, _, _ in deeplabv2_param_shapes(num_classes, output_stride, blocks)]
        self.mult = lr_multipliers(self.names)
        self.base_lr, self.momentum, self.wd = lr, momentum, weight_decay
        self.max_iters, self.power = max_iters, power
        self.cur_iter = 1            # _LRScheduler.__init__ already stepped once (lrer.py:152)
        self.cons_scale, self.rampup_steps, self.ema_decay = cons_scale, rampup_steps, ema_decay
        self.cons_for_labeled = cons_for_labeled
        self.os, self.blocks, self.ignore = output_stride, blocks, ignore_index
        self.bufs = [torch.zeros_like(self.s[n]) for n in self.names]
        self.step_idx = 0

    def step(self, img, gt, lbs, t_img=None):
        """One training step; returns dict of the scalars / tensors the goldens record."""
        for n in self.names:
            self.s[n].requires_grad_(True)
            self.s[n].grad = None
        s_logits, _ = deeplabv2_forward(img, self.s, True, self.os, self.blocks)
        s_task = sseg_criterion(s_logits[:lbs], gt[:lbs], self.ignore).mean()
        out = {'s_logits': s_logits.detach(), 's_task_loss': s_task.detach()}
        loss = s_task
        if self.t is not None:
            with torch.no_grad():
                t_logits, _ = deeplabv2_forward(img if t_img is None else t_img, self.t, True,
                                                self.os, self.blocks)
                out['t_task_loss'] = sseg_criterion(t_logits[:lbs], gt[:lbs], self.ignore).mean()
                out['t_logits'] = t_logits
            ramp = sigmoid_rampup(self.step_idx, self.rampup_steps)
            if self.cons_for_labeled:
                cons = mse_consistency(s_logits, t_logits)
            elif img.shape[0] > lbs:
                cons = mse_consistency(s_logits[lbs:], t_logits[lbs:])
            else:
                cons = torch.zeros(())
            cons = ramp * self.cons_scale * cons
            out['cons_loss'] = cons.detach()
            loss = loss + cons
        loss.backward()
        grads = [self.s[n].grad for n in self.names]
        out['grads'] = {n: g.detach().clone() for n, g in zip(self.names, grads)}
        lrs = [poly_lr(self.base_lr * m, self.cur_iter, self.max_iters, self.power) for m in self.mult]
        with torch.no_grad():
            for n in self.names:
                self.s[n].requires_grad_(False)
            sgd_momentum_step([self.s[n] for n in self.names], grads, self.bufs, lrs,
                              self.momentum, self.wd, first_step=(self.step_idx == 0))
            if self.t is not None:
                ema_update([self.t[n] for n in self.names], [self.s[n] for n in self.names],
                           self.ema_decay, self.step_idx)
        self.cur_iter += 1
        self.step_idx += 1
        return out


class CutMixOracle(MTOracle):
    """Functional restatement of ``SSLCUTMIX._train``'s loop body (ssl_cutmix.py:140-251) and
    ``_batch_prehandle`` (:383-432): labeled rows -> student -> CE; unlabeled rows -> teacher
    (no grad) -> softmax -> halves mixed with the box mask -> one confidence scalar; mixed
    images -> student -> MSE(softmax, mixed pseudo label) * confidence * rampup * cons_scale."""

    def __init__(self, *a, cons_threshold=0.97, mask_prop_range=(0.5, 0.5), **k):
        super().__init__(*a, **k)
        self.thr, self.prop = cons_threshold, mask_prop_range

    def step(self, img, gt, lbs, rng):
        ubs = img.shape[0] - lbs
        half = ubs // 2
        for n in self.names:
            self.s[n].requires_grad_(True)
            self.s[n].grad = None
        masks, _ = box_masks(rng, half, tuple(img.shape[2:]), prop_range=self.prop)
        mask = torch.from_numpy(masks).to(img.dtype)
        mix_inp = cutmix_mix(mask, img[lbs:lbs + half], img[lbs + half:lbs + ubs])
        l_logits, _ = deeplabv2_forward(img[:lbs], self.s, True, self.os, self.blocks)
        task = sseg_criterion(l_logits, gt[:lbs], self.ignore).mean()
        with torch.no_grad():
            t_logits, _ = deeplabv2_forward(img[lbs:lbs + ubs], self.t, True, self.os, self.blocks)
            t_prob = channel_softmax(t_logits)
            mp = cutmix_mix(mask, t_prob[:half], t_prob[half:ubs])
            conf = cutmix_confidence(mp, self.thr).to(img.dtype)
        u_logits, _ = deeplabv2_forward(mix_inp, self.s, True, self.os, self.blocks)
        ramp = sigmoid_rampup(self.step_idx, self.rampup_steps)
        cons = ramp * self.cons_scale * (F.mse_loss(channel_softmax(u_logits), mp) * conf)
        (task + cons).backward()
        grads = [self.s[n].grad for n in self.names]
        out = {'task_loss': task.detach(), 'cons_loss': cons.detach(), 'confidence': conf,
               'grads': {n: g.detach().clone() for n, g in zip(self.names, grads)}}
        lrs = [poly_lr(self.base_lr * m, self.cur_iter, self.max_iters, self.power) for m in self.mult]
        with torch.no_grad():
            for n in self.names:
                self.s[n].requires_grad_(False)
            sgd_momentum_step([self.s[n] for n in self.names], grads, self.bufs, lrs,
                              self.momentum, self.wd, first_step=(self.step_idx == 0))
            ema_update([self.t[n] for n in self.names], [self.s[n] for n in self.names],
                       self.ema_decay, self.step_idx)
        self.cur_iter += 1
        self.step_idx += 1
        return out


def to_dtype(state, dtype):
    """Cast a model state (floating tensors only) - used for the fp64 'exact arithmetic' runs that
    measure how far the reference's own fp32 evaluation is from the true value."""
    return {k: (v.to(dtype) if v.is_floating_point() else v.clone()) for k, v in state.items()}


def synthetic_batch(seed, batch, lbs, h, w, num_classes=21, ignore_frac=0.05, ignore_index=255):
    """SURVEY.md 8(d) synthetic inputs: randn images (ImageNet-normalised look-alike,
    task/sseg/data.py:99), integer-valued float labels with ~5% ignore pixels on the labeled
    rows, -1 on the unlabeled rows (data.py:105), labeled rows first (nn/data.py:156-159)."""
    g = torch.Generator().manual_seed(seed)
    img = torch.randn(batch, 3, h, w, generator=g)
    lab = torch.randint(0, num_classes, (batch, 1, h, w), generator=g).float()
    ign = torch.rand(batch, 1, h, w, generator=g) < ignore_frac
    lab[ign] = float(ignore_index)
    lab[lbs:] = -1.0
    return img, lab


# ----------------------------------------------------------------------------------------------
# PSPNet (task/sseg/module/_pspnet.py)
# ----------------------------------------------------------------------------------------------

R50_BLOCKS = (3, 4, 6, 3)
PSP_BINS = (1, 2, 3, 6)


def pspnet_param_shapes(num_classes=21, output_stride=16, blocks=R50_BLOCKS):
    """(name, shape, kind) in ``_PSPNet.parameters()`` order: backbone, psp (4 stages + bottleneck), decoder."""
    out = [(n, s, k) for n, s, k in deeplabv2_param_shapes(num_classes, output_stride, blocks) if n.startswith('backbone.')]
    for i in range(4):
        out += [('psp.stages.%d.1.weight' % i, (512, 2048, 1, 1), 'psp_conv'),
                ('psp.stages.%d.2.weight' % i, (512,), 'bn_w'), ('psp.stages.%d.2.bias' % i, (512,), 'bn_b')]
    out += [('psp.bottleneck.0.weight', (512, 4096, 3, 3), 'psp_conv'),
            ('psp.bottleneck.1.weight', (512,), 'bn_w'), ('psp.bottleneck.1.bias', (512,), 'bn_b'),
            ('decoder.0.weight', (num_classes, 512, 1, 1), 'dec_conv')]
    for j in (1, 2, 3):
        out += [('decoder.%d.conv.weight' % j, (num_classes * 4, num_classes, 1, 1), 'dec_conv'),
                ('decoder.%d.conv.bias' % j, (num_classes * 4,), 'dec_bias')]
    return out


def init_pspnet(seed, num_classes=21, output_stride=16, blocks=R50_BLOCKS):
    g = torch.Generator().manual_seed(seed)
    st = {}
    for name, shape, kind in pspnet_param_shapes(num_classes, output_stride, blocks):
        if kind == 'conv':
            st[name] = torch.randn(shape, generator=g) * math.sqrt(2.0 / (shape[2] * shape[3] * shape[0]))
        elif kind in ('psp_conv', 'dec_conv'):
            st[name] = torch.randn(shape, generator=g) * math.sqrt(2.0 / (shape[1] * shape[2] * shape[3]))
        elif kind == 'dec_bias':
            st[name] = (torch.rand(shape, generator=g) * 2 - 1) * 0.05
        elif kind == 'bn_w':
            st[name] = torch.ones(shape)
        else:
            st[name] = torch.zeros(shape)
    bn_names = [n[:-len('.weight')] for n, s, k in pspnet_param_shapes(num_classes, output_stride, blocks) if k == 'bn_w']
    for n in bn_names:
        c = st[n + '.weight'].numel()
        st[n + '.running_mean'] = torch.zeros(c)
        st[n + '.running_var'] = torch.ones(c)
        st[n + '.num_batches_tracked'] = torch.tensor(0, dtype=torch.long)
    return st


def pspnet_forward(img, state, training=True, output_stride=16, blocks=R50_BLOCKS):
    """_PSPNet.forward (_pspnet.py:122-128) + _PSPModule.forward (:96-102) + upsample / PixelShuffle (:15-54)."""
    bx = resnet_forward(img, state, training, output_stride, blocks)
    h, w = bx.shape[2:]
    pyramids = [bx]
    for i, b in enumerate(PSP_BINS):
        y = F.adaptive_avg_pool2d(bx, b)
        y = F.conv2d(y, state['psp.stages.%d.1.weight' % i])
        y = F.relu(batch_norm(y, state, 'psp.stages.%d.2' % i, training))
        pyramids.append(F.interpolate(y, size=(h, w), mode='bilinear', align_corners=False))
    x = F.conv2d(torch.cat(pyramids, dim=1), state['psp.bottleneck.0.weight'], padding=1)
    px = F.relu(batch_norm(x, state, 'psp.bottleneck.1', training))
    x = F.conv2d(px, state['decoder.0.weight'])
    for j in (1, 2, 3):
        x = F.pixel_shuffle(F.relu(F.conv2d(x, state['decoder.%d.conv.weight' % j], state['decoder.%d.conv.bias' % j])), 2)
    return F.interpolate(x, size=img.shape[2:], mode='bilinear', align_corners=True), px


# ----------------------------------------------------------------------------------------------
# validation metrics and the MT input-noise layer
# ----------------------------------------------------------------------------------------------
def confusion_matrix(pred, gt, num_classes):
    """SemanticSegmentationFunc.metrics, task/sseg/func.py:39-47: argmax over channels, pixels with
    0 <= gt < C, bincount of C*gt + pred.  Rows = ground truth."""
    pred = np.asarray(pred)
    gt = np.asarray(gt)
    arg = np.expand_dims(np.argmax(pred, axis=1), axis=1)
    mask = (gt >= 0) & (gt < num_classes)
    label = num_classes * gt[mask].astype('int') + arg[mask]
    return np.bincount(label, minlength=num_classes ** 2).reshape(num_classes, num_classes)


def seg_metrics(cmat_sum):
    """acc, acc-class, mIoU, fwIoU of the accumulated confusion matrix, task/sseg/func.py:64-80."""
    cmat_sum = np.asarray(cmat_sum)
    with np.errstate(divide='ignore', invalid='ignore'):
        acc = np.diag(cmat_sum).sum() / cmat_sum.sum()
        acc_class = np.nanmean(np.diag(cmat_sum) / cmat_sum.sum(axis=1))
        iou = np.diag(cmat_sum) / (np.sum(cmat_sum, axis=1) + np.sum(cmat_sum, axis=0) - np.diag(cmat_sum))
        miou = np.nanmean(iou)
        freq = np.sum(cmat_sum, axis=1) / np.sum(cmat_sum)
        fwiou = (freq[freq > 0] * iou[freq > 0]).sum()
    return np.array([acc, acc_class, miou, fwiou])


def gaussian_noise_layer(inp, noise):
    """GaussianNoiseLayer.forward, pixelssl/nn/module/gaussian_noise.py:18-41, for a given noise tensor
    (the layer draws it as N(0, uniform(0, std))).  Out of place."""
    x = inp.clone()
    imax = x.amax(dim=(1, 2, 3), keepdim=True)
    imin = x.amin(dim=(1, 2, 3), keepdim=True)
    rng = imax - imin + 1e-9
    x = (x - imin) / rng
    x = x + noise
    upper = (x > 1.0).float()
    lower = (x < 0.0).float()
    x = x * (1 - upper) + upper
    x = x * (1 - lower)
    return x * rng + imin
