"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the patch transform of the reference's feature-extraction loop
(datasets/dataset_h5.py:20-37 ``eval_transforms``: transforms.Resize(224) on a PIL RGB image, ToTensor, Normalize;
applied at dataset_h5.py:229 and consumed by Step2_feature_extract.py:58-66).

The resize is Pillow's 8-bit antialiased BILINEAR resample (Pillow src/libImaging/Resample.c: ``precompute_coeffs``,
``normalize_coeffs_8bpc``, horizontal pass then vertical pass, each rounded to uint8); Pillow (12.2 here, an unpinned
dependency of torchvision in requirements.txt) is a third-party dependency whose source is not under /root/reference, so
its published algorithm is restated.  Parity status: PINNED -- tests/test_preprocess_oracle.py runs the very torchvision /
PIL call sequence of the reference on random patches in the build container and requires byte equality of the resample and
bit equality of the normalised floats.
"""
from __future__ import annotations

import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2
IMAGENET_MEAN = (0.485, 0.456, 0.406)      # dataset_h5.py:22-23 (pretrained=True)
IMAGENET_STD = (0.229, 0.224, 0.225)


def precompute_coeffs(in_size: int, out_size: int):
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int64)
    kk = np.zeros((out_size, ksize), np.int64)
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        ss = 1.0 / filterscale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = np.array([max(0.0, 1.0 - abs((x + xmin - center + 0.5) * ss)) for x in range(xmax)])
        ww = w.sum()
        if ww != 0.0:
            w = w / ww
        kk[xx, :xmax] = [int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS)) for v in w]
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _pass(img, bounds, kk, axis):
    """One 8-bit resample pass along `axis` of an [H, W, C] uint8 image."""
    src = np.moveaxis(img.astype(np.int64), axis, 0)
    out = np.empty((bounds.shape[0],) + src.shape[1:], np.int64)
    for xx in range(bounds.shape[0]):
        xmin, cnt = bounds[xx]
        acc = (1 << (PRECISION_BITS - 1)) + np.tensordot(kk[xx, :cnt], src[xmin:xmin + cnt], axes=(0, 0))
        out[xx] = np.clip(acc >> PRECISION_BITS, 0, 255)
    return np.moveaxis(out, 0, axis).astype(np.uint8)


def resize_bilinear_u8(img: np.ndarray, out_size: int) -> np.ndarray:
    """PIL ``Image.resize((out, out), BILINEAR)`` of an [H, W, 3] uint8 array: horizontal pass, then vertical."""
    h, w, _ = img.shape
    t = _pass(img, *precompute_coeffs(w, out_size), axis=1) if w != out_size else img
    return _pass(t, *precompute_coeffs(h, out_size), axis=0) if h != out_size else t


def eval_transform(patches: np.ndarray, out_size: int = 224, mean=IMAGENET_MEAN, std=IMAGENET_STD) -> np.ndarray:
    """[B, H, W, 3] uint8 -> [B, 3, out, out] float32 = Normalize(ToTensor(Resize(out)(patch)))."""
    out = np.empty((patches.shape[0], 3, out_size, out_size), np.float32)
    m = np.asarray(mean, np.float32)[:, None, None]
    s = np.asarray(std, np.float32)[:, None, None]
    for i, pimg in enumerate(patches):
        r = resize_bilinear_u8(pimg, out_size).transpose(2, 0, 1).astype(np.float32)
        out[i] = (r / np.float32(255.0) - m) / s
    return out
