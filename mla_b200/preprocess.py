"""GPU image preprocessing: the reference's CPU-side `CLIPImageProcessor(672)` + ones mask channel
(vla/datasets/datasets.py:53-69; models/mla/model_mla.py:661-665; models/mla/image/vision_tokenizer.py:98-105) on
uint8 camera frames, bit-exact (csrc/preprocess.cu).

Host side here: PIL's bicubic coefficient tables (Pillow libImaging/Resample.c `precompute_coeffs` +
`normalize_coeffs_8bpc`: float64 weights normalised per output pixel, rounded to 22-bit fixed point) and the 3 x 256
uint8 -> normalised-f32 lookup table in transformers-4.40.1 arithmetic (`rescale`: float64 product cast to float32;
`normalize`: float32 subtract and divide by the CLIP statistics).  Both depend only on the frame size, so they are
built once and cached on the device.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, Tuple

import numpy as np
import torch

from . import _lib, ops
from ._lib import check

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)
_BITS = 22


def _cubic(x: float) -> float:
    x = abs(x)
    if x < 1.0:
        return (1.5 * x - 2.5) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * -0.5
    return 0.0


def bicubic_table(in_size: int, out_size: int) -> np.ndarray:
    """int32 [out_size, 2 + ksize]: first input index, tap count, fixed-point taps (zero padded)."""
    scale = in_size / out_size
    fs = max(scale, 1.0)
    support = 2.0 * fs
    ksize = int(math.ceil(support)) * 2 + 1
    tab = np.zeros((out_size, 2 + ksize), np.int32)
    for o in range(out_size):
        center = (o + 0.5) * scale
        lo = max(int(center - support + 0.5), 0)
        hi = min(int(center + support + 0.5), in_size)
        w = [_cubic((i + lo - center + 0.5) / fs) for i in range(hi - lo)]
        tot = 0.0
        for v in w:
            tot += v
        if tot != 0.0:
            w = [v / tot for v in w]
        tab[o, 0], tab[o, 1] = lo, hi - lo
        for i, v in enumerate(w):
            tab[o, 2 + i] = int(v * (1 << _BITS) - 0.5) if v < 0 else int(v * (1 << _BITS) + 0.5)
    return tab


def normalise_lut() -> np.ndarray:
    """f32 [3, 256]: (float32(float64(u) * (1/255)) - mean_c) / std_c in float32."""
    u = np.arange(256, dtype=np.float64)
    x = (u * (1 / 255)).astype(np.float32)
    mean, std = np.array(CLIP_MEAN, np.float32), np.array(CLIP_STD, np.float32)
    return np.ascontiguousarray(((x[:, None] - mean) / std).T.astype(np.float32))


_CACHE: Dict[Tuple, Tuple[torch.Tensor, torch.Tensor, torch.Tensor, int, int]] = {}


def _tables(h: int, w: int, size: int, device):
    key = (h, w, size, str(device))
    if key not in _CACHE:
        th, tv = bicubic_table(w, size), bicubic_table(h, size)
        _CACHE[key] = (torch.from_numpy(th).to(device), torch.from_numpy(tv).to(device),
                       torch.from_numpy(normalise_lut()).to(device), th.shape[1] - 2, tv.shape[1] - 2)
    return _CACHE[key]


def _frames(frames: torch.Tensor) -> torch.Tensor:
    if frames.dtype != torch.uint8 or frames.dim() != 4 or frames.shape[-1] != 3:
        raise _lib.MlaError(f"expected uint8 camera frames [B, H, W, 3], got {frames.dtype} {tuple(frames.shape)}")
    if not frames.is_cuda:
        raise _lib.MlaError("frames must be on the GPU (libmla_b200 has no CPU path)")
    if frames.shape[1] != frames.shape[2]:
        raise _lib.MlaError("shortest-edge resize + centre crop is built for square frames (all of the reference's cameras)")
    return frames.contiguous()


def clip_preprocess(frames: torch.Tensor, size: int = 672, add_mask: bool = True) -> torch.Tensor:
    """uint8 [B, H, W, 3] -> f32 [B, 4 (or 3), size, size]: exactly the tensor the reference's collator carries."""
    f = _frames(frames)
    B, H, W, _ = f.shape
    th, tv, lut, kh, kv = _tables(H, W, size, f.device)
    ch = 4 if add_mask else 3
    out = torch.empty((B, ch, size, size), dtype=torch.float32, device=f.device)
    check(_lib.lib().mla_clip_preprocess(ops._p(f), ops._p(th), ops._p(tv), ops._p(lut), ops._p(out), C.c_int32(B),
                                         C.c_int32(H), C.c_int32(W), C.c_int32(size), C.c_int32(kh), C.c_int32(kv),
                                         C.c_int32(ch), ops._stream()))
    return out


def patchify_frames(frames: torch.Tensor, size: int, patch: int, conv_stride: int, k_pad: int) -> torch.Tensor:
    """uint8 [B, H, W, 3] -> bf16 im2col rows [B*(size/patch)^2, k_pad] of the patch embedding, window-major — the
    rows mla_patchify produces from clip_preprocess(frames), without materialising that tensor."""
    f = _frames(frames)
    B, H, W, _ = f.shape
    th, tv, lut, kh, kv = _tables(H, W, size, f.device)
    rows = B * (size // patch) ** 2
    out = torch.empty((rows, k_pad), dtype=torch.bfloat16, device=f.device)
    check(_lib.lib().mla_patchify_u8(ops._p(f), ops._p(th), ops._p(tv), ops._p(lut), ops._p(out), C.c_int32(B),
                                     C.c_int32(H), C.c_int32(W), C.c_int32(size), C.c_int32(kh), C.c_int32(kv),
                                     C.c_int32(patch), C.c_int32(conv_stride), C.c_int32(k_pad), ops._stream()))
    return out
