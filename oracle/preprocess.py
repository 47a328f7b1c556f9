"""Oracle (test infrastructure): the reference's image preprocessing restated in numpy integer / float arithmetic.

The data side (vla/datasets/datasets.py:53-69, RLDSBatchTransform) and inference (models/mla/model_mla.py:661-665) turn
a uint8 HxWx3 camera frame (224x224 for RLBench) into the model's f32 [4, 672, 672] input with
`CLIPImageProcessor(do_resize=True, size=672, do_center_crop=True, crop_size=672, do_rescale=True, do_normalize=True)`
(models/mla/image/vision_tokenizer.py:98-105) + an all-ones mask channel:
  1. PIL `Image.resize(..., BICUBIC)` — third-party (Pillow; libImaging/Resample.c): separable convolution on uint8 with
     22-bit fixed-point coefficients, horizontal pass then vertical pass, each rounded and clamped to uint8.  Restated
     here from Pillow's published algorithm and pinned BIT-EXACT against the Pillow in this image
     (tests/test_preprocess_cpu.py);
  2. rescale: uint8 -> float64 * (1/255) -> float32;   3. normalise: (x - mean) / std in float32 (CLIP statistics);
  4. concat a ones channel.  Pinned against transformers' CLIPImageProcessor in the same test.
"""
from __future__ import annotations

import math

import numpy as np

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)
PRECISION_BITS = 32 - 8 - 2


def _bicubic(x: float) -> float:
    a = -0.5
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def resample_coeffs(in_size: int, out_size: int):
    """precompute_coeffs + normalize_coeffs_8bpc of Pillow's Resample.c for the bicubic filter over the full image.
    Returns (xmin int32 [out], xcount int32 [out], k int32 [out, ksize])."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    xmin = np.zeros(out_size, np.int32)
    xcnt = np.zeros(out_size, np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        lo = max(int(center - support + 0.5), 0)
        hi = min(int(center + support + 0.5), in_size)
        n = hi - lo
        w = [_bicubic((x + lo - center + 0.5) * ss) for x in range(n)]
        ww = sum(w)          # sequential float64 sum, like the C loop
        if ww != 0.0:
            w = [v / ww for v in w]
        xmin[xx], xcnt[xx] = lo, n
        for x, v in enumerate(w):
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
    return xmin, xcnt, kk


def _pass(img: np.ndarray, out_size: int, axis: int) -> np.ndarray:
    """One separable pass along `axis` (0 = vertical, 1 = horizontal) on uint8 [H, W, C]."""
    xmin, xcnt, kk = resample_coeffs(img.shape[axis], out_size)
    src = np.moveaxis(img, axis, 0).astype(np.int64)                  # [in, other, C]
    out = np.empty((out_size,) + src.shape[1:], np.uint8)
    for xx in range(out_size):
        n, lo = int(xcnt[xx]), int(xmin[xx])
        acc = np.full(src.shape[1:], 1 << (PRECISION_BITS - 1), np.int64)
        for x in range(n):
            acc += src[lo + x] * int(kk[xx, x])
        out[xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def pil_resize_bicubic(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """PIL.Image.resize((out_w, out_h), BICUBIC) of a uint8 [H, W, C] array: horizontal pass, then vertical pass."""
    assert img.dtype == np.uint8 and img.ndim == 3
    if img.shape[1] != out_w:
        img = _pass(img, out_w, 1)
    if img.shape[0] != out_h:
        img = _pass(img, out_h, 0)
    return img


def clip_preprocess(img: np.ndarray, size: int = 672, add_mask: bool = True) -> np.ndarray:
    """uint8 [H, W, 3] (square, as every camera of the reference's datasets) -> f32 [4, size, size]."""
    assert img.shape[0] == img.shape[1], "shortest-edge resize + centre crop restated for square frames only"
    r = pil_resize_bicubic(img, size, size)
    x = (r.astype(np.float64) * (1 / 255)).astype(np.float32)
    x = (x - np.array(CLIP_MEAN, np.float32)) / np.array(CLIP_STD, np.float32)
    x = np.ascontiguousarray(x.transpose(2, 0, 1))
    if add_mask:
        x = np.concatenate([x, np.ones((1, size, size), np.float32)], 0)
    return x
