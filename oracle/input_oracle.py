"""CPU oracle of the sseg TRAINING INPUT PIPELINE (task/sseg/data.py:90-123, 142-256) - groundwork for SURVEY.md
section 8(f) rank 4 (GPU augmentation).  TEST INFRASTRUCTURE ONLY, same rules as oracle/sseg_oracle.py.

The reference builds every training sample on the host with PIL: random scale of the short edge (BILINEAR for the
image, NEAREST for the label), zero padding up to the crop size, random crop, random horizontal flip, ImageNet
normalisation, HWC -> CHW float32.  The arithmetic that matters lives in an un-vendored dependency (Pillow, unpinned
in pixelssl/requirements.txt; Pillow 12.2 in this image): ``Image.resize`` on 8-bit images is a separable,
antialiased (support scaled by the reduction factor) convolution in 22-bit fixed point with an 8-bit intermediate
(libImaging/Resample.c), NEAREST is ``floor`` of a source coordinate that starts at scale/2 and is incremented by scale
(libImaging/Geometry.c, ImagingScaleAffine).  Both are restated here in
numpy and pinned bit for bit against Pillow itself and against the reference's transform classes
(tests/test_oracle_golden.py); the random draws follow the reference's order on Python's ``random`` module."""
import random as _random

import numpy as np

PRECISION_BITS = 32 - 8 - 2          # libImaging/Resample.c
MEAN = (0.485, 0.456, 0.406)         # task/sseg/data.py:98,112
STD = (0.229, 0.224, 0.225)


def _bilinear_filter(x):
    x = abs(x)
    return 1.0 - x if x < 1.0 else 0.0


def _coefficients(in_size, out_size):
    """precompute_coeffs + normalize_coeffs_8bpc of Resample.c for the bilinear filter (support 1.0) over the whole
    input range: per output sample the first input index, the tap count and the 22-bit fixed-point weights."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ss = 1.0 / filterscale
    bounds, taps = [], []
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        k = [_bilinear_filter((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = sum(k)
        if ww != 0.0:
            k = [v / ww for v in k]
        fixed = [int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS)) for v in k]
        bounds.append((xmin, xmax))
        taps.append(np.array(fixed, dtype=np.int64))
    return bounds, taps


def _resample_axis0(a, out_size):
    """One pass of ImagingResampleVertical_8bpc along axis 0 of a uint8 array [n, ...]."""
    bounds, taps = _coefficients(a.shape[0], out_size)
    out = np.empty((out_size,) + a.shape[1:], dtype=np.uint8)
    src = a.astype(np.int64)
    for xx, ((xmin, xmax), k) in enumerate(zip(bounds, taps)):
        acc = np.full(a.shape[1:], 1 << (PRECISION_BITS - 1), dtype=np.int64)
        for x in range(xmax):
            acc += src[xmin + x] * k[x]
        out[xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return out


def resize_bilinear_u8(img, ow, oh):
    """``Image.resize((ow, oh), Image.BILINEAR)`` for a uint8 image [H, W] or [H, W, C]: horizontal pass first, then
    vertical, each rounding to 8 bits (ImagingResample)."""
    a = np.ascontiguousarray(img)
    if a.dtype != np.uint8:
        raise TypeError('8-bit images only')
    h, w = a.shape[:2]
    if ow != w:
        a = np.swapaxes(_resample_axis0(np.swapaxes(a, 0, 1), ow), 0, 1)
    if oh != h:
        a = _resample_axis0(a, oh)
    return np.ascontiguousarray(a)


def resize_nearest(img, ow, oh):
    """``Image.resize((ow, oh), Image.NEAREST)`` (ImagingScaleAffine)."""
    a = np.asarray(img)
    h, w = a.shape[:2]

    def table(n_in, n_out):
        # the source coordinate is ACCUMULATED in double precision (xo += scale), not recomputed per sample
        step = n_in / n_out
        xo, idx = step * 0.5, []
        for _ in range(n_out):
            idx.append(min(int(xo), n_in - 1))
            xo += step
        return np.array(idx, dtype=np.int64)
    return np.ascontiguousarray(a[table(h, oh)][:, table(w, ow)])


def random_scale_crop(img, mask, base_size, crop_size, rng=_random, fill=0):
    """RandomScaleCrop.__call__ (task/sseg/data.py:223-256): three draws - short edge, crop x, crop y."""
    short_size = rng.randint(int(base_size * 0.5), int(base_size * 2.0))
    h, w = img.shape[:2]
    if h > w:
        ow = short_size
        oh = int(1.0 * h * ow / w)
    else:
        oh = short_size
        ow = int(1.0 * w * oh / h)
    img = resize_bilinear_u8(img, ow, oh)
    mask = resize_nearest(mask, ow, oh)
    if short_size < crop_size:
        padh = crop_size - oh if oh < crop_size else 0
        padw = crop_size - ow if ow < crop_size else 0
        img = np.pad(img, ((0, padh), (0, padw)) + ((0, 0),) * (img.ndim - 2), constant_values=0)
        mask = np.pad(mask, ((0, padh), (0, padw)) + ((0, 0),) * (mask.ndim - 2), constant_values=fill)
    h, w = img.shape[:2]
    x1 = rng.randint(0, w - crop_size)
    y1 = rng.randint(0, h - crop_size)
    return img[y1:y1 + crop_size, x1:x1 + crop_size], mask[y1:y1 + crop_size, x1:x1 + crop_size]


def random_horizontal_flip(img, mask, rng=_random):
    """RandomHorizontalFlip.__call__ (data.py:184-192): one draw."""
    if rng.random() < 0.5:
        return img[:, ::-1], mask[:, ::-1]
    return img, mask


def normalize_to_chw(img, mask):
    """Normalize + ToTensor (data.py:142-181): float32 arithmetic in the reference's order, HWC -> CHW."""
    x = np.array(img).astype(np.float32)
    x /= 255.0
    x -= MEAN
    x /= STD
    return np.ascontiguousarray(x.astype(np.float32).transpose(2, 0, 1)), np.array(mask).astype(np.float32)


def train_prehandle(image, label, base_size, crop_size, rng=_random):
    """PascalVocDataset._train_prehandle (data.py:90-109).  ``label`` None (unlabeled sample): the image stands in for
    the label through the geometric transforms and the returned label is the constant -1 map (data.py:105)."""
    lab = image if label is None else label
    img, lab = random_scale_crop(image, lab, base_size, crop_size, rng)
    img, lab = random_horizontal_flip(img, lab, rng)
    x, y = normalize_to_chw(img, lab)
    if label is None:
        return x, x[0] * 0.0 - 1.0
    return x, y


def fixed_scale_resize(img, mask, size):
    """FixedScaleResize.__call__ (data.py:259-292): short edge -> ``size`` (float ratio, truncated), BILINEAR / NEAREST,
    then zero padding on the right / bottom up to ``size`` (cv2.copyMakeBorder, constant 0)."""
    h, w = img.shape[:2]
    if w <= h:
        ow = size
        oh = h * ow / w
    else:
        oh = size
        ow = w * oh / h
    oh, ow = int(oh), int(ow)
    img = resize_bilinear_u8(img, ow, oh)
    mask = resize_nearest(mask, ow, oh)
    pad_w, pad_h = max(size - ow, 0), max(size - oh, 0)
    if pad_w > 0 or pad_h > 0:
        img = np.pad(img, ((0, pad_h), (0, pad_w)) + ((0, 0),) * (img.ndim - 2), constant_values=0)
        mask = np.pad(mask, ((0, pad_h), (0, pad_w)), constant_values=0)
    return img, mask


def val_prehandle(image, label, im_size=None, rescaling=False):
    """PascalVocDataset._val_prehandle (data.py:111-125)."""
    if rescaling:
        image, label = fixed_scale_resize(image, label, im_size)
    return normalize_to_chw(image, label)
