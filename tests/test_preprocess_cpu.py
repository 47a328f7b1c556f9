"""Pins the image-preprocessing oracle (oracle/preprocess.py) bit-exactly: against the golden recorded from the
reference's vendored CLIPImageProcessor (tests/golden/make_golden_preprocess.py), against the Pillow in this image
for the resize alone, and the product's host-side tables (mla_b200/preprocess.py) against the oracle's."""
import hashlib
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "preprocess.npz")


def test_oracle_reproduces_reference_preprocessing_bit_exact():
    from oracle import preprocess as P
    z = np.load(GOLD)
    for name in ("a", "b"):
        out = P.clip_preprocess(z[f"{name}.frame"], add_mask=False)
        assert out.dtype == np.float32 and out.shape == (3, 672, 672)
        assert np.array_equal(out[:, ::37, ::41], z[f"{name}.probe"])
        digest = np.frombuffer(hashlib.sha256(np.ascontiguousarray(out).tobytes()).digest(), np.uint8)
        assert np.array_equal(digest, z[f"{name}.sha256"]), name
    full = P.clip_preprocess(z["a.frame"])
    assert full.shape == (4, 672, 672) and bool((full[3] == 1).all())


def test_oracle_resize_equals_pillow():
    Image = pytest.importorskip("PIL.Image")
    from oracle import preprocess as P
    rs = np.random.RandomState(3)
    for hw, out in ((224, 672), (97, 672), (672, 672), (900, 672), (224, 300)):
        a = (rs.rand(hw, hw, 3) * 255).astype(np.uint8)
        a[:5, :5], a[-5:, -5:] = 255, 0
        ref = np.asarray(Image.fromarray(a).resize((out, out), resample=Image.BICUBIC))
        assert np.array_equal(P.pil_resize_bicubic(a, out, out), ref), (hw, out)


def test_product_tables_equal_oracle_tables():
    """Host logic of the CUDA path: PIL's fixed-point taps and the uint8 -> normalised-f32 table."""
    from mla_b200 import preprocess as Q
    from oracle import preprocess as P
    for i, o in ((224, 672), (180, 672), (97, 672), (672, 672), (900, 672)):
        xmin, xcnt, kk = P.resample_coeffs(i, o)
        t = Q.bicubic_table(i, o)
        assert t.dtype == np.int32 and t.shape == (o, 2 + kk.shape[1])
        assert np.array_equal(t[:, 0], xmin) and np.array_equal(t[:, 1], xcnt) and np.array_equal(t[:, 2:], kk)
        assert (t[:, 2:].sum(1) - (1 << 22)).__abs__().max() <= 4          # taps sum to 1.0 in fixed point
    u = np.arange(256, dtype=np.uint8).reshape(16, 16, 1).repeat(3, 2)
    x = (u.astype(np.float64) * (1 / 255)).astype(np.float32)
    x = (x - np.array(P.CLIP_MEAN, np.float32)) / np.array(P.CLIP_STD, np.float32)
    assert np.array_equal(Q.normalise_lut().T.reshape(16, 16, 3), x)
