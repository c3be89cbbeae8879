"""CPU oracle of the CCT step (pixelssl/ssl_algorithm/ssl_cct.py).  TEST INFRASTRUCTURE ONLY (see
oracle/sseg_oracle.py).  Pinned by tests/golden/cct_step_65.npz (oracle/make_golden.py:golden_cct)."""
import math
import random

import numpy as np
import torch
import torch.nn.functional as F

from . import sseg_oracle as O

KINDS = ('vat', 'drop', 'cut', 'context', 'object', 'fd', 'fn')


def decoder_param_shapes(idx, in_channels=2048, nc=21):
    """Parameters of one auxiliary decoder's ``upsample`` (ssl_cct.py:531-539) in module order."""
    p = 'auxiliary_decoders.%d.upsample.' % idx
    out = [(p + '0.weight', (nc, in_channels, 1, 1))]
    for j in (1, 2, 3):
        out += [(p + '%d.conv.weight' % j, (nc * 4, nc, 1, 1)), (p + '%d.conv.bias' % j, (nc * 4,))]
    return out


def init_decoders(seed, n, in_channels=2048, nc=21):
    g = torch.Generator().manual_seed(seed)
    st = {}
    for i in range(n):
        for name, shape in decoder_param_shapes(i, in_channels, nc):
            if name.endswith('weight'):
                st[name] = torch.randn(shape, generator=g) * math.sqrt(2.0 / shape[1])
            else:
                st[name] = (torch.rand(shape, generator=g) * 2 - 1) * 0.05
    return st


def upsample_forward(st, idx, x):
    """upsample(): conv1x1 (no bias) then 3 x [conv1x1 (bias) -> ReLU -> PixelShuffle(2)] (ssl_cct.py:501-539)."""
    p = 'auxiliary_decoders.%d.upsample.' % idx
    x = F.conv2d(x, st[p + '0.weight'])
    for j in (1, 2, 3):
        x = F.pixel_shuffle(F.relu(F.conv2d(x, st[p + '%d.conv.weight' % j], st[p + '%d.conv.bias' % j])), 2)
    return x


def _l2_normalize(d):
    return d / (torch.norm(d.reshape(d.shape[0], -1), dim=1).view(-1, 1, 1, 1) + 1e-8)


def vat_r_adv(st, idx, x, xi, eps):
    """VATDecoder.get_r_adv, ssl_cct.py:555-576 (one iteration)."""
    xd = x.detach()
    with torch.no_grad():
        pred = F.softmax(upsample_forward(st, idx, xd), dim=1)
    d = _l2_normalize(torch.rand(x.shape).sub(0.5).to(x.dtype))
    d.requires_grad_()
    logp_hat = F.log_softmax(upsample_forward(st, idx, xd + xi * d), dim=1)
    adv = F.kl_div(logp_hat, pred, reduction='batchmean')
    (g,) = torch.autograd.grad(adv, d)
    return _l2_normalize(g) * eps


def guided_cutout(output, resize, erase):
    """CutOutDecoder.guided_cutout, ssl_cct.py:604-651."""
    import cv2
    masks = (output.argmax(1) > 0).float()
    outs = []
    for mask in masks:
        mask_np = np.uint8(mask.numpy())
        ones = np.ones_like(mask_np)
        found = cv2.findContours(mask_np, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
        contours = found[0] if len(found) == 2 else found[1]
        for c in contours:
            if c.shape[0] <= 50:
                continue
            poly = c.reshape(c.shape[0], c.shape[-1])
            min_w, max_w, min_h, max_h = poly[:, 0].min(), poly[:, 0].max(), poly[:, 1].min(), poly[:, 1].max()
            bb_w, bb_h = max_w - min_w, max_h - min_h
            sw = random.randint(0, int(bb_w * (1 - erase)))
            sh = random.randint(0, int(bb_h * (1 - erase)))
            ones[min_h + sh:min_h + sh + int(bb_h * erase), min_w + sw:min_w + sw + int(bb_w * erase)] = 0
        outs.append(ones)
    m = torch.from_numpy(np.stack(outs)).float().unsqueeze(1)
    return F.interpolate(m, size=resize, mode='nearest')


def decoder_forward(st, idx, kind, x, main_pred, cfg):
    """The seven auxiliary decoders (ssl_cct.py:542-745); draws come from the global python / numpy /
    torch CPU generators exactly like the reference."""
    size = (x.size(2), x.size(3))
    if kind == 'vat':
        x = x + vat_r_adv(st, idx, x, cfg['xi'], cfg['eps']).to(x.dtype)
    elif kind == 'drop':
        x = F.dropout2d(x, cfg['drop_rate'], training=True)
    elif kind == 'cut':
        x = x * guided_cutout(main_pred, size, cfg['erase']).to(x.dtype)
    elif kind in ('context', 'object'):
        m = F.interpolate((main_pred.argmax(1) > 0).float().unsqueeze(1), size=size, mode='nearest').to(x.dtype)
        x = x * (m if kind == 'context' else 1 - m)
    elif kind == 'fd':
        att = torch.mean(x, dim=1, keepdim=True)
        mx = att.view(x.size(0), -1).max(dim=1, keepdim=True)[0]
        thr = (mx * np.random.uniform(0.7, 0.9)).view(x.size(0), 1, 1, 1).expand_as(att)
        x = x.mul((att < thr).to(x.dtype))
    elif kind == 'fn':
        noise = torch.distributions.uniform.Uniform(-cfg['uniform'], cfg['uniform']).sample(x.shape[1:]).unsqueeze(0)
        x = x.mul(noise.to(x.dtype)) + x
    return upsample_forward(st, idx, x)


class CctOracle(O.MTOracle):
    """SSLCCT._train loop body (ssl_cct.py:226-301) + WrappedCCTModel.forward (:438-491)."""

    def __init__(self, s_state, dec_state, kinds, cons_scale=30.0, rampup_steps=0, ad_lr_scale=10.0, xi=1e-6, eps=2.0,
                 drop_rate=0.5, erase=0.4, uniform=0.3, **k):
        super().__init__(s_state, None, **k)
        self.dec, self.kinds = dec_state, list(kinds)
        self.dec_names = [n for i in range(len(kinds)) for n, _ in decoder_param_shapes(i)]
        self.cfg = {'xi': xi, 'eps': eps, 'drop_rate': drop_rate, 'erase': erase, 'uniform': uniform}
        self.cons_scale, self.rampup_steps, self.ad_lr_scale = cons_scale, rampup_steps, ad_lr_scale
        self.dec_bufs = [torch.zeros_like(self.dec[n]) for n in self.dec_names]

    def step(self, img, gt, lbs):
        for n in self.names:
            self.s[n].requires_grad_(True)
            self.s[n].grad = None
        for n in self.dec_names:
            self.dec[n].requires_grad_(True)
            self.dec[n].grad = None
        l_logits, _ = O.deeplabv2_forward(img[:lbs], self.s, True, self.os, self.blocks)
        task = O.sseg_criterion(l_logits, gt[:lbs], self.ignore).mean()
        u_logits, latent = O.deeplabv2_forward(img[lbs:], self.s, True, self.os, self.blocks)
        target = O.channel_softmax(u_logits).detach()
        main_pred = u_logits.detach()
        cons = 0
        for i, kind in enumerate(self.kinds):
            low = decoder_forward(self.dec, i, kind, latent, main_pred, self.cfg)
            up = F.interpolate(low, size=target.shape[2:], mode='bilinear')
            cons = cons + F.mse_loss(F.softmax(up, dim=1), target)
        cons = cons / len(self.kinds)
        ramp = O.sigmoid_rampup(self.step_idx, self.rampup_steps)
        cons = ramp * self.cons_scale * cons
        (task + cons).backward()
        out = {'task_loss': task.detach(), 'cons_loss': cons.detach(),
               'grads': {n: self.s[n].grad.detach().clone() for n in self.names},
               'dec_grads': {n: self.dec[n].grad.detach().clone() for n in self.dec_names}}
        lrs = [O.poly_lr(self.base_lr * m, self.cur_iter, self.max_iters, self.power) for m in self.mult]
        dlr = O.poly_lr(self.base_lr * self.ad_lr_scale, self.cur_iter, self.max_iters, self.power)
        with torch.no_grad():
            for n in self.names:
                self.s[n].requires_grad_(False)
            for n in self.dec_names:
                self.dec[n].requires_grad_(False)
            O.sgd_momentum_step([self.s[n] for n in self.names], [out['grads'][n] for n in self.names], self.bufs, lrs,
                                self.momentum, self.wd, first_step=(self.step_idx == 0))
            O.sgd_momentum_step([self.dec[n] for n in self.dec_names], [out['dec_grads'][n] for n in self.dec_names],
                                self.dec_bufs, [dlr] * len(self.dec_names), self.momentum, self.wd,
                                first_step=(self.step_idx == 0))
        self.cur_iter += 1
        self.step_idx += 1
        return out
