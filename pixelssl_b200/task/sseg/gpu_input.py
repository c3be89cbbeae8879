"""GPU input pipeline of the sseg task: ``PascalVocDataset._train_prehandle`` / ``_val_prehandle``
(task/sseg/data.py:90-123) with the pixel work in ONE CUDA launch per sample (csrc/input_pipeline.cu) instead of
PIL + numpy on the host, bit for bit the same tensors.

The host keeps what is inherently sequential and tiny: the random draws (Python's ``random`` module, in the reference's
order: short edge, crop x, crop y - RandomScaleCrop, data.py:229-251 - then the flip coin, data.py:187) and the
per-axis resampling tables of Pillow's ``Image.resize`` (libImaging/Resample.c precompute_coeffs /
normalize_coeffs_8bpc for BILINEAR, libImaging/Geometry.c ImagingScaleAffine for NEAREST), a few hundred integers
per sample.  The 8-bit source image goes to HBM as it is (3 B/pixel instead of the 12 B/pixel float tensor the
reference uploads) and the normalised CHW crop is produced where the model reads it."""
import ctypes
import random as _random

import numpy as np
import torch

from ... import ops
from ..._lib import call

PRECISION_BITS = 32 - 8 - 2          # libImaging/Resample.c
MEAN = (0.485, 0.456, 0.406)         # task/sseg/data.py:98,112
STD = (0.229, 0.224, 0.225)
_MEAN_C = (ctypes.c_double * 3)(*MEAN)
_STD_C = (ctypes.c_double * 3)(*STD)


def bilinear_tables(in_size, out_size):
    """Pillow's antialiased BILINEAR resampling along one axis: -> (bounds int32 [out,2] = first source index and tap
    count, weights int32 [out,kmax] in 22-bit fixed point).  Double arithmetic in Pillow's order so that the rounded
    integers are identical (precompute_coeffs + normalize_coeffs_8bpc, support 1.0 scaled by the reduction factor)."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ss = 1.0 / filterscale
    kmax = int(np.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    weights = np.zeros((out_size, kmax), dtype=np.int32)
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        k = []
        for x in range(xmax):
            a = abs((x + xmin - center + 0.5) * ss)
            k.append(1.0 - a if a < 1.0 else 0.0)
        ww = sum(k)
        if ww != 0.0:
            k = [v / ww for v in k]
        bounds[xx] = (xmin, xmax)
        for x, v in enumerate(k):
            weights[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
    return bounds, weights


def nearest_table(n_in, n_out):
    """Image.resize(NEAREST): the source coordinate is accumulated in double precision (xo += scale)."""
    step = n_in / n_out
    xo = step * 0.5
    idx = np.empty(n_out, dtype=np.int32)
    for i in range(n_out):
        idx[i] = min(int(xo), n_in - 1)
        xo += step
    return idx


def draw_train_geometry(h, w, base_size, crop_size, rng=_random):
    """The four random draws of RandomScaleCrop + RandomHorizontalFlip in the reference's order
    -> (ow, oh, x1, y1, flip)."""
    short_size = rng.randint(int(base_size * 0.5), int(base_size * 2.0))
    if h > w:
        ow = short_size
        oh = int(1.0 * h * ow / w)
    else:
        oh = short_size
        ow = int(1.0 * w * oh / h)
    pw, ph = ow, oh
    if short_size < crop_size:
        ph = max(oh, crop_size)
        pw = max(ow, crop_size)
    x1 = rng.randint(0, pw - crop_size)
    y1 = rng.randint(0, ph - crop_size)
    flip = rng.random() < 0.5
    return ow, oh, x1, y1, flip


def _dev(a, device):
    return torch.from_numpy(np.ascontiguousarray(a)).to(device, non_blocking=True)


def _launch(img, lab, ow, oh, x1, y1, crop_w, crop_h, flip, label_const, device):
    h, w = img.shape[:2]
    img_d = _dev(img, device) if not torch.is_tensor(img) else img
    lab_d = None if lab is None else (_dev(lab, device) if not torch.is_tensor(lab) else lab)
    no_resize = int(ow == w and oh == h)
    P = ops._p
    if no_resize:
        tabs = [None] * 6
        kx = ky = 0
    else:
        xb, xk = bilinear_tables(w, ow)
        yb, yk = bilinear_tables(h, oh)
        kx, ky = xk.shape[1], yk.shape[1]
        tabs = [_dev(t, device) for t in (xb, xk, yb, yk, nearest_table(w, ow), nearest_table(h, oh))]
    out_img = torch.empty((3, crop_h, crop_w), dtype=torch.float32, device=device)
    out_lab = torch.empty((crop_h, crop_w), dtype=torch.float32, device=device)
    call('pxl_input_prehandle', P(img_d), P(lab_d), h, w, ow, oh, no_resize, P(tabs[0]), P(tabs[1]), kx, P(tabs[2]), P(tabs[3]), ky,
         P(tabs[4]), P(tabs[5]), int(x1), int(y1), int(crop_w), int(crop_h), int(bool(flip)), 0.0, float(label_const),
         _MEAN_C, _STD_C, P(out_img), P(out_lab), ops._stream())
    return out_img, out_lab


def train_prehandle(image, label, base_size, crop_size, rng=_random, device='cuda'):
    """PascalVocDataset._train_prehandle (data.py:90-109) for one sample.  ``image``: uint8 [H,W,3] (numpy or a
    device tensor), ``label``: uint8 [H,W] or None (unlabeled: the returned label is the constant -1 map, data.py:105).
    -> (float32 [3,crop,crop], float32 [crop,crop]) on ``device``."""
    h, w = image.shape[:2]
    ow, oh, x1, y1, flip = draw_train_geometry(h, w, base_size, crop_size, rng)
    return _launch(image, label, ow, oh, x1, y1, crop_size, crop_size, flip, -1.0, device)


def val_prehandle(image, label, im_size=None, rescaling=False, device='cuda'):
    """PascalVocDataset._val_prehandle (data.py:111-125): optional FixedScaleResize (data.py:259-292: short edge ->
    im_size with a float ratio, BILINEAR / NEAREST, zero padding right / below up to im_size), Normalize, ToTensor."""
    h, w = image.shape[:2]
    if not rescaling:
        return _launch(image, label, w, h, 0, 0, w, h, False, 0.0, device)
    if w <= h:
        ow = im_size
        oh = h * ow / w
    else:
        oh = im_size
        ow = w * oh / h
    oh, ow = int(oh), int(ow)
    return _launch(image, label, ow, oh, 0, 0, max(ow, im_size), max(oh, im_size), False, 0.0, device)
