"""Generates tests/golden/preprocess.npz from the reference's own image path: the vendored transformers 4.40.1
`CLIPImageProcessor(do_resize=True, size=672, do_center_crop=True, crop_size=672, do_normalize=True, do_rescale=True)`
(models/mla/image/vision_tokenizer.py:98-105) applied to a PIL image, as vla/datasets/datasets.py:53-56 does.

    python tests/golden/make_golden_preprocess.py     # needs /root/reference

Stored: two uint8 frames (224x224: the RLBench camera size; 180x180: a non-integer resize ratio), and for each the
SHA-256 of the f32 [3,672,672] result plus a strided probe of its values (the full tensors would be 5.4 MB each).
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def frame(hw, seed):
    rs = np.random.RandomState(seed)
    yy, xx = np.mgrid[0:hw, 0:hw].astype(np.float32) / hw
    base = np.stack([yy, xx, 1 - yy * xx], -1) * 200 + rs.rand(hw, hw, 3) * 55        # smooth ramp + texture
    img = base.astype(np.uint8)
    img[: hw // 8, : hw // 8] = 255                                                    # saturated blocks: the bicubic
    img[-hw // 8:, -hw // 8:] = 0                                                      # overshoot must clamp like PIL
    return img


def main():
    ref_shim.load()
    import transformers
    assert transformers.__version__ == "4.40.1", transformers.__version__
    from PIL import Image
    from transformers import CLIPImageProcessor
    ip = CLIPImageProcessor(do_resize=True, size=672, do_center_crop=True, crop_size=672, do_normalize=True,
                            do_rescale=True)
    save = {}
    for name, hw, seed in (("a", 224, 0), ("b", 180, 1)):
        img = frame(hw, seed)
        out = ip.preprocess(Image.fromarray(img), return_tensors="pt")["pixel_values"][0].numpy()
        assert out.dtype == np.float32 and out.shape == (3, 672, 672)
        save[f"{name}.frame"] = img
        save[f"{name}.sha256"] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(out).tobytes()).digest(), np.uint8)
        save[f"{name}.probe"] = out[:, ::37, ::41].copy()
    np.savez_compressed(os.path.join(OUT, "preprocess.npz"), **save)
    print("wrote", os.path.getsize(os.path.join(OUT, "preprocess.npz")), "bytes")


if __name__ == "__main__":
    main()
