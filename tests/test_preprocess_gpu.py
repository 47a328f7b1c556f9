"""GPU image preprocessing (csrc/preprocess.cu) against the oracle — bit-exact (integer resize, table-driven
normalisation) — and the fused preprocess+im2col path against the two-step path."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "preprocess.npz")


@pytest.mark.parametrize("hw", [224, 180, 97, 672, 900])
def test_clip_preprocess_bit_exact(cuda_lib, hw):
    from mla_b200.preprocess import clip_preprocess
    from oracle import preprocess as P
    z = np.load(GOLD)
    if hw in (224, 180):
        frames = np.stack([z["a.frame" if hw == 224 else "b.frame"]] * 2)
        frames[1] = frames[1][::-1, :, ::-1]
    else:
        rs = np.random.RandomState(hw)
        frames = (rs.rand(2, hw, hw, 3) * 255).astype(np.uint8)
        frames[:, :9, :9], frames[:, -9:, -9:] = 255, 0
    got = clip_preprocess(torch.from_numpy(np.ascontiguousarray(frames)).cuda()).cpu().numpy()
    assert got.shape == (2, 4, 672, 672) and got.dtype == np.float32
    for b in range(2):
        want = P.clip_preprocess(np.ascontiguousarray(frames[b]))
        assert np.array_equal(got[b], want), (hw, b, np.abs(got[b] - want).max())


def test_fused_patchify_equals_two_step(cuda_lib):
    """bf16 im2col rows straight from uint8 frames == mla_patchify(clip_preprocess(frames)), bit for bit."""
    from mla_b200 import _lib, ops
    from mla_b200.preprocess import clip_preprocess, patchify_frames
    z = np.load(GOLD)
    frames = torch.from_numpy(np.stack([z["a.frame"], z["a.frame"][:, ::-1].copy()])).cuda()
    px = clip_preprocess(frames)
    B, P_, cs, k_pad = 2, 14, 3, 592
    cols = torch.empty((B * 48 * 48, k_pad), dtype=torch.bfloat16, device="cuda")
    _lib.check(cuda_lib.mla_patchify(ops._p(px), ops._p(cols), C.c_int32(B), C.c_int32(4), C.c_int32(672), C.c_int32(672),
                                     C.c_int32(P_), C.c_int32(cs), C.c_int32(k_pad), ops._stream()))
    fused = patchify_frames(frames, 672, P_, cs, k_pad)
    assert torch.equal(fused, cols)


def test_tokenizer_accepts_raw_frames(cuda_lib):
    """VisionTokenizer on uint8 camera frames == on the reference's preprocessed f32 tensor (same pooled features)."""
    from mla_b200.preprocess import clip_preprocess
    from mla_b200.vision import VisionTokenizer
    torch.manual_seed(0)
    vt = VisionTokenizer(1024).cuda().eval().requires_grad_(False)
    z = np.load(GOLD)
    frames = torch.from_numpy(np.stack([z["a.frame"]])).cuda()
    a, h, w = vt.pooled_features(frames)
    b, h2, w2 = vt.pooled_features(clip_preprocess(frames))
    assert (h, w) == (h2, w2) == (16, 16)
    assert torch.equal(a, b)


def test_bad_inputs_raise(cuda_lib):
    from mla_b200 import _lib
    from mla_b200.preprocess import clip_preprocess
    with pytest.raises(_lib.MlaError):
        clip_preprocess(torch.zeros(1, 224, 224, 3, device="cuda"))                      # not uint8
    with pytest.raises(_lib.MlaError):
        clip_preprocess(torch.zeros(1, 224, 200, 3, dtype=torch.uint8, device="cuda"))   # not square
    with pytest.raises(_lib.MlaError):
        clip_preprocess(torch.zeros(1, 224, 224, 3, dtype=torch.uint8))                  # host tensor
