"""CPU oracle of the GCT step (pixelssl/ssl_algorithm/ssl_gct.py).  TEST INFRASTRUCTURE ONLY (see
oracle/sseg_oracle.py).  Pinned by tests/golden/gct_step_129.npz (oracle/make_golden.py:golden_gct)."""
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import sseg_oracle as O

FD_SPEC = [('conv1', 'ibn1', None, 64, 2), ('conv2', 'ibn2', 64, 128, 2), ('conv2_1', 'ibn2_1', 128, 128, 1),
           ('conv3', 'ibn3', 128, 256, 2), ('conv3_1', 'ibn3_1', 256, 256, 1), ('conv4', 'ibn4', 256, 512, 2),
           ('conv4_1', 'ibn4_1', 512, 512, 1)]


def fd_param_shapes(in_channels=24):
    """FlawDetector parameters in module order (ssl_gct.py:549-563)."""
    out = []
    for cname, iname, cin, cout, stride in FD_SPEC:
        cin = in_channels if cin is None else cin
        nb = int(cout * 0.5 + 0.5)
        out += [(cname + '.weight', (cout, cin, 4, 4)), (cname + '.bias', (cout,)),
                (iname + '.bnorm.weight', (nb,)), (iname + '.bnorm.bias', (nb,))]
    out += [('classifier.weight', (1, 512, 4, 4)), ('classifier.bias', (1,))]
    return out


def init_fd(seed, in_channels=24, classifier_gain=40.0):
    """classifier_gain scales the last conv so random-init flaw maps exceed FlawmapHandler's 0.1 clip
    threshold (otherwise every golden would exercise only the 'map zeroed' branch)."""
    g = torch.Generator().manual_seed(seed)
    st = {}
    for name, shape in fd_param_shapes(in_channels):
        if name.endswith('.weight') and len(shape) == 4:
            st[name] = (torch.rand(shape, generator=g) * 2 - 1) / math.sqrt(shape[1] * 16)
        elif 'bnorm.weight' in name:
            st[name] = 1.0 + 0.2 * torch.randn(shape, generator=g)
        elif 'bnorm.bias' in name:
            st[name] = 0.1 * torch.randn(shape, generator=g)
        else:
            st[name] = (torch.rand(shape, generator=g) * 2 - 1) * 0.05
    st['classifier.weight'] = st['classifier.weight'] * classifier_gain
    for cname, iname, cin, cout, stride in FD_SPEC:
        nb = int(cout * 0.5 + 0.5)
        st[iname + '.bnorm.running_mean'] = torch.zeros(nb)
        st[iname + '.bnorm.running_var'] = torch.ones(nb)
        st[iname + '.bnorm.num_batches_tracked'] = torch.tensor(0, dtype=torch.long)
    return st


def ibnorm(x, st, prefix, training=True):
    """IBNorm.forward, ssl_gct.py:600-607."""
    nb = st[prefix + '.bnorm.weight'].numel()
    xb = F.batch_norm(x[:, :nb].contiguous(), st[prefix + '.bnorm.running_mean'], st[prefix + '.bnorm.running_var'],
                      st[prefix + '.bnorm.weight'], st[prefix + '.bnorm.bias'], training, 0.1, 1e-5)
    xi = F.instance_norm(x[:, nb:].contiguous(), eps=1e-5)
    return torch.cat((xb, xi), 1)


def fd_forward(st, task_inp, task_pred, training=True):
    """FlawDetector.forward, ssl_gct.py:565-585."""
    x = torch.cat((task_inp, task_pred), dim=1)
    for cname, iname, cin, cout, stride in FD_SPEC:
        x = F.conv2d(x, st[cname + '.weight'], st[cname + '.bias'], stride=stride, padding=1)
        x = F.leaky_relu(ibnorm(x, st, iname, training), 0.2)
    x = F.conv2d(x, st['classifier.weight'], st['classifier.bias'], stride=2, padding=1)
    return F.interpolate(x, size=task_pred.shape[2:], mode='bilinear', align_corners=True)


def odd_ksize(v):
    k = int(v)
    return k + 1 if k % 2 == 0 else k


def flawmap_handler(flawmap, im_size, clip_threshold=0.1):
    """FlawmapHandler.forward, ssl_gct.py:641-657 - including the in-place clamp of the input's data."""
    fm = flawmap.data
    fm.mul_((fm >= 0).to(fm.dtype))
    fm = O.gaussian_blur(fm, odd_ksize(im_size / 16)) if fm.dtype == torch.float32 else _blur_any(fm, odd_ksize(im_size / 16))
    fmax = fm.amax(dim=(1, 2, 3), keepdim=True)
    fmin = fm.amin(dim=(1, 2, 3), keepdim=True)
    fm = fm * (fmax > clip_threshold).to(fm.dtype)
    return (fm - fmin) / (fmax - fmin + 1e-9)


def _blur_any(x, k):
    w = torch.from_numpy(O.gaussian_kernel_2d(k)).to(x.dtype)[None, None]
    return F.conv2d(F.pad(x, (k // 2,) * 4, mode='reflect'), w)


def dcgt(l_pred, r_pred, l_handled, r_handled, thr):
    """DCGTGenerator.forward, ssl_gct.py:668-689."""
    l_tmp, r_tmp = l_handled.clone(), r_handled.clone()
    both_bad = ((l_tmp > thr) & (r_tmp > thr)).to(l_pred.dtype)
    lh = l_handled * (l_tmp <= thr).to(l_pred.dtype) + (l_tmp > thr).to(l_pred.dtype)
    rh = r_handled * (r_tmp <= thr).to(l_pred.dtype) + (r_tmp > thr).to(l_pred.dtype)
    l_mask = (rh >= lh).to(l_pred.dtype)
    r_mask = (lh >= rh).to(l_pred.dtype)
    return l_mask * l_pred + (1 - l_mask) * r_pred, r_mask * r_pred + (1 - r_mask) * l_pred, both_bad


def prepare_gt_for_fdgt(task_gt, num_classes=21, ignore_index=255):
    """sslgct_prepare_task_gt_for_fdgt, task/sseg/func.py:179-192."""
    keep = (task_gt != ignore_index).to(task_gt.dtype)
    return torch.cat([(task_gt == i).to(task_gt.dtype) * keep for i in range(num_classes)], dim=1)


def fdgt(pred, gt_onehot, im_size, mu, nu):
    """FDGTGenerator.forward, ssl_gct.py:714-728."""
    blur = O.gaussian_blur if pred.dtype == torch.float32 else _blur_any
    diff = (gt_onehot - pred.detach()).abs().sum(dim=1, keepdim=True) * mu
    diff = blur(diff, odd_ksize(im_size / 8))
    for _ in range(nu):
        dil = F.max_pool2d(F.pad(diff, (1, 1, 1, 1), mode='reflect'), 3, stride=1)
        diff = blur(dil, odd_ksize(im_size / 4))
    dmax = diff.amax(dim=(1, 2, 3), keepdim=True)
    dmin = diff.amin(dim=(1, 2, 3), keepdim=True)
    return (diff - dmin) / (dmax - dmin + 1e-9)


class GctOracle:
    """SSLGCT._train loop body (ssl_gct.py:185-293) with DeepLabV2 task models."""

    def __init__(self, l_state, r_state, fd_state, im_size, fc_ssl_scale=1.0, dc_ssl_scale=100.0, dc_threshold=0.6,
                 rampup_steps=0, fd_lr=1e-4, fd_scale=10.0, mu=0.5, nu=1, lr=0.00025, momentum=0.9, weight_decay=5e-4,
                 max_iters=10):
        self.l = O.MTOracle(l_state, None, lr=lr, momentum=momentum, weight_decay=weight_decay, max_iters=max_iters)
        self.r = O.MTOracle(r_state, None, lr=lr, momentum=momentum, weight_decay=weight_decay, max_iters=max_iters)
        self.fd = fd_state
        self.fd_names = [n for n, _ in fd_param_shapes()]
        for n in self.fd_names:
            self.fd[n].requires_grad_(True)
        self.fd_opt = torch.optim.Adam([self.fd[n] for n in self.fd_names], lr=fd_lr, betas=(0.9, 0.99))
        self.im, self.fc, self.dc, self.thr = im_size, fc_ssl_scale, dc_ssl_scale, dc_threshold
        self.rampup_steps, self.fd_lr, self.fd_scale, self.mu, self.nu = rampup_steps, fd_lr, fd_scale, mu, nu
        self.max_iters, self.step_idx = max_iters, 0

    def _model_iter(self, m, img, gt, lbs, dc_gt, fc_mask, ramp, out, mid):
        for n in m.names:
            m.s[n].requires_grad_(True)
            m.s[n].grad = None
        for n in self.fd_names:
            self.fd[n].requires_grad_(False)
        logits, _ = O.deeplabv2_forward(img, m.s, True)
        prob = O.channel_softmax(logits)
        flawmap = fd_forward(self.fd, img, prob)
        task = O.sseg_criterion(logits[:lbs], gt[:lbs]).mean()
        fc = self.fc * (fc_mask * F.mse_loss(flawmap, torch.zeros_like(flawmap), reduction='none')).mean()
        dc = ramp * self.dc * F.mse_loss(prob, dc_gt)
        (task + fc + dc).backward()
        grads = [m.s[n].grad for n in m.names]
        out[mid + '_task_loss'], out[mid + '_fc_loss'], out[mid + '_dc_loss'] = task.detach(), fc.detach(), dc.detach()
        out[mid + '_grads'] = {n: g.detach().clone() for n, g in zip(m.names, grads)}
        lrs = [O.poly_lr(m.base_lr * k, m.cur_iter, m.max_iters, m.power) for k in m.mult]
        with torch.no_grad():
            for n in m.names:
                m.s[n].requires_grad_(False)
            O.sgd_momentum_step([m.s[n] for n in m.names], grads, m.bufs, lrs, m.momentum, m.wd, first_step=(m.step_idx == 0))
        m.cur_iter += 1
        m.step_idx += 1

    def step(self, img, gt, lbs):
        out = {}
        ramp = O.sigmoid_rampup(self.step_idx, self.rampup_steps)
        with torch.no_grad():
            l_act = O.channel_softmax(O.deeplabv2_forward(img, self.l.s, True)[0])
            r_act = O.channel_softmax(O.deeplabv2_forward(img, self.r.s, True)[0])
        for n in self.fd_names:
            self.fd[n].requires_grad_(True)
        l_flawmap = fd_forward(self.fd, img, l_act)
        r_flawmap = fd_forward(self.fd, img, r_act)
        with torch.no_grad():
            l_h = flawmap_handler(l_flawmap, self.im)
            r_h = flawmap_handler(r_flawmap, self.im)
            l_dc_gt, r_dc_gt, both_bad = dcgt(l_act, r_act, l_h, r_h, self.thr)
        out['both_bad_frac'] = both_bad.mean()
        self._model_iter(self.l, img, gt, lbs, l_dc_gt, both_bad, ramp, out, 'l')
        self._model_iter(self.r, img, gt, lbs, r_dc_gt, both_bad, ramp, out, 'r')
        for n in self.fd_names:
            self.fd[n].requires_grad_(True)
        with torch.no_grad():
            l_fm_gt = fdgt(l_act[:lbs], prepare_gt_for_fdgt(gt[:lbs]), self.im, self.mu, self.nu)
            r_fm_gt = fdgt(r_act[:lbs], prepare_gt_for_fdgt(gt[:lbs]), self.im, self.mu, self.nu)
        l_fd = self.fd_scale * F.mse_loss(l_flawmap[:lbs], l_fm_gt, reduction='none').mean(dim=(1, 2, 3)).mean()
        r_fd = self.fd_scale * F.mse_loss(r_flawmap[:lbs], r_fm_gt, reduction='none').mean(dim=(1, 2, 3)).mean()
        self.fd_opt.zero_grad()
        ((l_fd + r_fd) / 2).backward()
        out['l_fd_loss'], out['r_fd_loss'] = l_fd.detach(), r_fd.detach()
        out['fd_grads'] = {n: self.fd[n].grad.detach().clone() for n in self.fd_names}
        for grp in self.fd_opt.param_groups:
            grp['lr'] = O.poly_lr(self.fd_lr, self.step_idx + 1, self.max_iters, 0.9)
        self.fd_opt.step()
        self.step_idx += 1
        return out
