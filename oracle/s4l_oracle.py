"""CPU oracle of the S4L step (pixelssl/ssl_algorithm/ssl_s4l.py).  TEST INFRASTRUCTURE ONLY - same rules as
oracle/sseg_oracle.py (never imported by the product).  Pinned by tests/golden/s4l_step_65.npz, generated from the
unmodified reference by oracle/make_golden.py."""
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import sseg_oracle as O


def rc_shapes(c=21):
    """RotationClassifer parameters / buffers in module order (ssl_s4l.py:382-391)."""
    return [('conv1.weight', (c, c, 4, 4)), ('conv1.bias', (c,)), ('bn1.weight', (c,)), ('bn1.bias', (c,)),
            ('conv2.weight', (2 * c, c, 4, 4)), ('conv2.bias', (2 * c,)), ('bn2.weight', (2 * c,)), ('bn2.bias', (2 * c,)),
            ('classifier.weight', (4, 2 * c)), ('classifier.bias', (4,))]


def init_rc(seed, c=21):
    """Deterministic rotation-classifier state incl. BN buffers (loaded into the reference for goldens)."""
    g = torch.Generator().manual_seed(seed)
    st = {}
    for name, shape in rc_shapes(c):
        if name.startswith('bn') and name.endswith('weight'):
            st[name] = 1.0 + 0.1 * (torch.rand(shape, generator=g) * 2 - 1)
        elif name.startswith('bn'):
            st[name] = 0.1 * (torch.rand(shape, generator=g) * 2 - 1)
        elif name.endswith('weight'):
            fan_in = int(np.prod(shape[1:]))
            st[name] = (torch.rand(shape, generator=g) * 2 - 1) / math.sqrt(fan_in)
        else:
            st[name] = (torch.rand(shape, generator=g) * 2 - 1) * 0.05
    for bn, ch in (('bn1', c), ('bn2', 2 * c)):
        st[bn + '.running_mean'] = torch.zeros(ch)
        st[bn + '.running_var'] = torch.ones(ch)
        st[bn + '.num_batches_tracked'] = torch.tensor(0, dtype=torch.long)
    return st


def rc_forward(st, task_pred, training=True):
    """RotationClassifer.forward, ssl_s4l.py:393-400 (nn.BatchNorm2d: momentum 0.1, eps 1e-5, per-batch statistics;
    running buffers updated in place like the module does)."""
    x = task_pred
    for conv, bn in (('conv1', 'bn1'), ('conv2', 'bn2')):
        x = F.conv2d(x, st[conv + '.weight'], st[conv + '.bias'], stride=2, padding=1)
        x = F.batch_norm(x, st[bn + '.running_mean'], st[bn + '.running_var'], st[bn + '.weight'], st[bn + '.bias'],
                         training, 0.1, 1e-5)
        if training:
            st[bn + '.num_batches_tracked'] += 1
        x = F.leaky_relu(x, 0.2)
    x = F.adaptive_avg_pool2d(x, (1, 1)).view(task_pred.shape[0], -1)
    return F.linear(x, st['classifier.weight'], st['classifier.bias'])


def rotate_tensor(t, angle_idx):
    """_rotate_tensor, ssl_s4l.py:352-360, on a [C,H,W] tensor."""
    if angle_idx == 1:
        return t.transpose(1, 2).flip(2)
    if angle_idx == 2:
        return t.flip(2).flip(1)
    if angle_idx == 3:
        return t.transpose(1, 2).flip(1)
    return t


def batch_prehandle(img, gt, angles):
    """_batch_prehandle (train), ssl_s4l.py:296-350: [bs,...] -> [2bs,...] + the rotation ground truth."""
    bs = img.shape[0]
    ri = torch.cat((img, torch.stack([rotate_tensor(img[i], int(angles[i])) for i in range(bs)])))
    rg = torch.cat((gt, torch.stack([rotate_tensor(gt[i], int(angles[i])) for i in range(bs)])))
    rot = torch.cat((torch.zeros(bs, dtype=torch.long), torch.as_tensor(np.asarray(angles), dtype=torch.long)))
    return ri, rg, rot


class S4LOracle(O.MTOracle):
    """SSLS4L._train loop body (ssl_s4l.py:120-175) for DeepLabV2 + RotationClassifer on CPU."""

    def __init__(self, s_state, rc_state, rotated_sup_scale=0.5, rotation_scale=1.0, **k):
        super().__init__(s_state, None, **k)
        self.rc = rc_state
        self.rc_names = [n for n, _ in rc_shapes()]
        self.rss, self.rs = rotated_sup_scale, rotation_scale
        self.rc_bufs = [torch.zeros_like(self.rc[n]) for n in self.rc_names]

    def step(self, img, gt, lbs, angles):
        """img / gt: the ORIGINAL batch [bs,...] (labeled first); angles: np array [bs] in {1,2,3}."""
        bs = img.shape[0]
        inp, g2, rot = batch_prehandle(img, gt, angles)
        for st, names in ((self.s, self.names), (self.rc, self.rc_names)):
            for n in names:
                st[n].requires_grad_(True)
                st[n].grad = None
        logits, _ = O.deeplabv2_forward(inp, self.s, True, self.os, self.blocks)
        pred_rot = rc_forward(self.rc, logits, True)
        unrot = O.sseg_criterion(logits[:lbs], g2[:lbs], self.ignore).mean()
        rotd = self.rss * O.sseg_criterion(logits[bs:bs + lbs], g2[bs:bs + lbs], self.ignore).mean()
        rloss = self.rs * F.cross_entropy(pred_rot, rot)
        (unrot + rotd + rloss).backward()
        out = {'unrotated_task_loss': unrot.detach(), 'rotated_task_loss': rotd.detach(), 'rotation_loss': rloss.detach(),
               'rotation_logits': pred_rot.detach()}
        grads = [self.s[n].grad for n in self.names]
        rgrads = [self.rc[n].grad for n in self.rc_names]
        out['grads'] = {n: g.detach().clone() for n, g in zip(self.names, grads)}
        out['rc_grads'] = {n: g.detach().clone() for n, g in zip(self.rc_names, rgrads)}
        lrs = [O.poly_lr(self.base_lr * m, self.cur_iter, self.max_iters, self.power) for m in self.mult]
        rlr = O.poly_lr(self.base_lr, self.cur_iter, self.max_iters, self.power)          # the rotation classifier's group: lr
        with torch.no_grad():
            for st, names in ((self.s, self.names), (self.rc, self.rc_names)):
                for n in names:
                    st[n].requires_grad_(False)
            O.sgd_momentum_step([self.s[n] for n in self.names], grads, self.bufs, lrs, self.momentum, self.wd,
                                first_step=(self.step_idx == 0))
            O.sgd_momentum_step([self.rc[n] for n in self.rc_names], rgrads, self.rc_bufs, [rlr] * len(rgrads), self.momentum,
                                self.wd, first_step=(self.step_idx == 0))
        acc = (pred_rot.detach().argmax(1) == rot).float().sum() * (100.0 / (2 * bs))
        out['rotation_acc'] = acc
        self.cur_iter += 1
        self.step_idx += 1
        return out
