"""Golden vectors for the image preprocessing row, generated in the authoring container with the reference's own
dependencies restated as it pins them: Pillow's `Image.resize(..., BICUBIC)` for the resize (the installed Pillow IS the
dependency), and the transformers-4.31 `CLIPImageProcessor.preprocess` sequence around it (the installed transformers 5.5 ships
a torchvision-backed processor whose bicubic differs from Pillow's by up to 2/255; the reference's `sample()` signature pins
transformers 4.31-4.33, whose processor calls Pillow).  `expand2square` is taken from the UNMODIFIED reference file when
/root/reference is present.

    python tests/golden/make_clip_preprocess_golden.py   ->  tests/golden/clip_preprocess.npz (+ .json with SHA-256 of big outputs)
"""
import hashlib
import importlib.util
import json
import os
import sys

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
MEAN = [0.48145466, 0.4578275, 0.40821073]
STD = [0.26862954, 0.26130258, 0.27577711]


def hf431_preprocess(pil_img, size):
    """transformers 4.31 CLIPImageProcessor.preprocess, operation by operation (image_processing_clip.py of that release):
    convert_to_rgb -> to_numpy_array -> resize(shortest_edge, PIL BICUBIC) -> center_crop -> rescale(1/255) -> normalize -> CHW."""
    img = pil_img.convert("RGB")
    w, h = img.size
    short, long = (w, h) if w <= h else (h, w)
    new_short, new_long = size, int(size * long / short)
    out_h, out_w = (new_long, new_short) if w <= h else (new_short, new_long)
    img = img.resize((out_w, out_h), resample=Image.BICUBIC)
    a = np.asarray(img)
    top, left = (out_h - size) // 2, (out_w - size) // 2
    a = a[top:top + size, left:left + size]
    x = (a * (1 / 255)).astype(np.float32)
    x = (x - np.array(MEAN, dtype=x.dtype)) / np.array(STD, dtype=x.dtype)
    return np.ascontiguousarray(x.transpose(2, 0, 1))


def reference_expand2square():
    path = "/root/reference/experiments/llava/mm_utils.py"
    if not os.path.exists(path):
        return None
    src = open(path).read()
    start = src.index("def expand2square")
    end = src.index("def process_images")
    ns = {"Image": Image}
    exec(compile(src[start:end], path, "exec"), ns)  # the unmodified function body
    return ns["expand2square"]


def synth_image(H, W, seed):
    """Deterministic test image (block structure + noise); the big cases are regenerated from (H, W, seed) instead of stored."""
    rng = np.random.default_rng(seed)
    smooth = rng.random((H // 8 + 2, W // 8 + 2, 3))
    img = np.kron(smooth, np.ones((8, 8, 1)))[:H, :W] * 200 + rng.random((H, W, 3)) * 55
    return img.astype(np.uint8)


def main():
    e2s = reference_expand2square()
    cases, arrays, meta = [], {}, {}
    # (name, H, W, size, pad)
    for name, H, W, S, pad in [("down_landscape_56", 120, 171, 56, False), ("down_portrait_56", 203, 97, 56, False),
                               ("up_56", 31, 45, 56, False), ("square_noresize_56", 56, 56, 56, False),
                               ("pad_landscape_56", 90, 160, 56, True), ("pad_portrait_56", 133, 71, 56, True),
                               ("coco_like_336", 480, 640, 336, False), ("pad_coco_like_336", 427, 640, 336, True)]:
        seed = 2024 + H * 7 + W
        img = synth_image(H, W, seed)
        pil = Image.fromarray(img)
        if pad:
            assert e2s is not None, "needs /root/reference for expand2square"
            pil = e2s(pil, tuple(int(x * 255) for x in MEAN))
        out = hf431_preprocess(pil, S)
        if S <= 64:
            arrays[name + ".image"] = img
            arrays[name + ".pixel_values"] = out
        meta[name] = {"H": H, "W": W, "size": S, "pad": pad, "seed": seed, "sha256": hashlib.sha256(out.tobytes()).hexdigest(),
                      "sum": float(out.astype(np.float64).sum())}
    np.savez_compressed(os.path.join(HERE, "clip_preprocess.npz"), **arrays)
    json.dump(meta, open(os.path.join(HERE, "clip_preprocess.json"), "w"), indent=1)
    print("wrote", len(meta), "cases;", os.path.getsize(os.path.join(HERE, "clip_preprocess.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
