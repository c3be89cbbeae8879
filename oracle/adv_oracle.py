"""CPU oracle of the AdvSSL step (pixelssl/ssl_algorithm/ssl_adv.py).  TEST INFRASTRUCTURE ONLY - same
rules as oracle/sseg_oracle.py (never imported by the product).  Pinned by
tests/golden/adv_step_65.npz, generated from the unmodified reference by oracle/make_golden.py."""
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import sseg_oracle as O

FCD_LAYERS = ('conv1', 'conv2', 'conv3', 'conv4', 'classifier')


def fcd_shapes(in_channels=21, ndf=64):
    """FCDiscriminator parameters in module order (ssl_adv.py:472-476)."""
    chans = [in_channels, ndf, ndf * 2, ndf * 4, ndf * 8, 1]
    out = []
    for i, name in enumerate(FCD_LAYERS):
        out += [(name + '.weight', (chans[i + 1], chans[i], 4, 4)), (name + '.bias', (chans[i + 1],))]
    return out


def init_fcd(seed, in_channels=21):
    """Deterministic discriminator state (nn.Conv2d-like scale; loaded into the reference for goldens)."""
    g = torch.Generator().manual_seed(seed)
    st = {}
    for name, shape in fcd_shapes(in_channels):
        if name.endswith('weight'):
            fan_in = shape[1] * 16
            st[name] = (torch.rand(shape, generator=g) * 2 - 1) / math.sqrt(fan_in)
        else:
            st[name] = (torch.rand(shape, generator=g) * 2 - 1) * 0.05
    return st


def fcd_forward(st, task_pred):
    """FCDiscriminator.forward, ssl_adv.py:480-493."""
    x = task_pred
    for name in FCD_LAYERS[:-1]:
        x = F.leaky_relu(F.conv2d(x, st[name + '.weight'], st[name + '.bias'], stride=2, padding=1), 0.2)
    x = F.conv2d(x, st['classifier.weight'], st['classifier.bias'], stride=2, padding=1)
    return F.interpolate(x, size=task_pred.shape[2:], mode='bilinear', align_corners=True)


def fcd_preprocess(fcd_pred, task_gt, is_real, ignore_index=255):
    """task/sseg/func.py:137-155: constant target, ignored pixels -> pred and target both zeroed."""
    biclass = 1.0 if is_real else 0.0
    if task_gt is None:
        ignore = torch.zeros_like(fcd_pred, dtype=torch.bool)
    else:
        ignore = task_gt == ignore_index
    fcd_gt = torch.full_like(fcd_pred, biclass)
    fcd_gt[ignore] = 255.0
    mask = ((fcd_gt >= 0) & (fcd_gt != ignore_index)).to(fcd_pred.dtype)
    return fcd_pred * mask, fcd_gt * mask


def fcd_criterion(pred, gt):
    """FCDiscriminatorCriterion, ssl_adv.py:496-503."""
    return F.binary_cross_entropy_with_logits(pred, gt, reduction='none').mean(dim=(1, 2, 3))


def onehot_gt(task_gt, num_classes=21):
    """ssladv_convert_task_gt_to_fcd_input, task/sseg/func.py:157-168."""
    return torch.cat([(task_gt == i).to(task_gt.dtype) for i in range(num_classes)], dim=1)


class AdvOracle(O.MTOracle):
    """SSLADV._train loop body (ssl_adv.py:126-279) for DeepLabV2 + FCDiscriminator on CPU."""

    def __init__(self, s_state, d_state, labeled_adv_scale=0.01, unlabeled_adv_scale=0.001, adv_for_labeled=True,
                 discriminator_lr=1e-4, discriminator_scale=1.0, unlabeled_for_discriminator=False,
                 discriminator_power=0.9, **k):
        super().__init__(s_state, None, **k)
        self.d = d_state
        self.d_names = [n for n, _ in fcd_shapes()]
        self.las, self.uas, self.afl = labeled_adv_scale, unlabeled_adv_scale, adv_for_labeled
        self.d_scale, self.ufd = discriminator_scale, unlabeled_for_discriminator
        self.d_base_lr, self.d_power = discriminator_lr, discriminator_power
        for n in self.d_names:
            self.d[n].requires_grad_(True)
        self.d_opt = torch.optim.Adam([self.d[n] for n in self.d_names], lr=discriminator_lr, betas=(0.9, 0.99))

    def step(self, img, gt, lbs):
        bs = img.shape[0]
        for n in self.names:
            self.s[n].requires_grad_(True)
            self.s[n].grad = None
        logits, _ = O.deeplabv2_forward(img, self.s, True, self.os, self.blocks)
        prob = O.channel_softmax(logits)
        conf = fcd_forward(self.d, prob)
        l_gt = gt[:lbs]
        task = O.sseg_criterion(logits[:lbs], l_gt, self.ignore).mean()
        out = {'task_loss': task.detach()}
        loss = task
        if self.afl:
            p, g = fcd_preprocess(conf[:lbs], l_gt, True, self.ignore)
            la = self.las * fcd_criterion(p, g).mean()
            out['labeled_adv_loss'] = la.detach()
            loss = loss + la
        if bs > lbs:
            p, g = fcd_preprocess(conf[lbs:], None, True, self.ignore)
            ua = self.uas * fcd_criterion(p, g).mean()
            out['unlabeled_adv_loss'] = ua.detach()
            loss = loss + ua
        loss.backward()
        grads = [self.s[n].grad for n in self.names]
        out['grads'] = {n: g.detach().clone() for n, g in zip(self.names, grads)}
        lrs = [O.poly_lr(self.base_lr * m, self.cur_iter, self.max_iters, self.power) for m in self.mult]
        with torch.no_grad():
            for n in self.names:
                self.s[n].requires_grad_(False)
            O.sgd_momentum_step([self.s[n] for n in self.names], grads, self.bufs, lrs, self.momentum, self.wd,
                                first_step=(self.step_idx == 0))
        # discriminator step
        self.d_opt.zero_grad()
        fake_pred = prob.detach() if self.ufd else prob[:lbs].detach()
        fconf = fcd_forward(self.d, fake_pred)
        p, g = fcd_preprocess(fconf[:lbs], l_gt, False, self.ignore)
        if self.ufd and bs > lbs:
            pu, gu = fcd_preprocess(fconf[lbs:], None, False, self.ignore)
            p, g = torch.cat((p, pu), 0), torch.cat((g, gu), 0)
        fake = self.d_scale * fcd_criterion(p, g).mean()
        rconf = fcd_forward(self.d, onehot_gt(l_gt))
        p, g = fcd_preprocess(rconf, l_gt, True, self.ignore)
        real = self.d_scale * fcd_criterion(p, g).mean()
        ((fake + real) / 2).backward()
        out['fake_d_loss'], out['real_d_loss'] = fake.detach(), real.detach()
        out['d_grads'] = {n: self.d[n].grad.detach().clone() for n in self.d_names}
        for grp in self.d_opt.param_groups:
            grp['lr'] = O.poly_lr(self.d_base_lr, self.cur_iter, self.max_iters, self.d_power)
        self.d_opt.step()
        self.cur_iter += 1
        self.step_idx += 1
        return out
