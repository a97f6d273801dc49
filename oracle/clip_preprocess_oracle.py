"""CPU oracle of the image preprocessing the reference's drivers run before `model.generate` (TEST INFRASTRUCTURE ONLY).

What it restates: `image_processor.preprocess(image, return_tensors='pt')['pixel_values'][0]` as called at
reference experiments/eval/sampling/llava_sampling.py:82 and eval/calibrate/llava_calibrate.py:149, with the image processor
that `load_pretrained_model` returns for LLaVA-1.5 (experiments/llava/model/builder.py:137-141 -> CLIPImageProcessor of
openai/clip-vit-large-patch14-336), and the optional `expand2square` padding of experiments/llava/mm_utils.py:152-164.

The arithmetic lives in un-vendored dependencies of the reference: transformers 4.31 `CLIPImageProcessor.preprocess`
(convert RGB -> resize shortest edge to 336 with PIL BICUBIC -> center crop 336 -> rescale 1/255 -> normalize by the CLIP
mean / std) and Pillow's `Image.resize` (ImagingResample: separable, anti-aliased, 8-bit fixed point with 22 fractional bits,
horizontal pass first, each pass rounded to uint8).  Both are restated here in numpy integer arithmetic.

Pin: tests/test_clip_preprocess.py checks `pil_resize_bicubic` BIT-EXACTLY against the installed Pillow's own
`Image.resize(..., Image.BICUBIC)` on random images (up- and down-scaling, odd sizes), and the whole pipeline against
the committed golden vectors tests/golden/clip_preprocess.npz, which tests/golden/make_clip_preprocess_golden.py generates with
Pillow + the transformers-4.31 operation sequence.
"""
from __future__ import annotations

import math

import numpy as np

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)
PRECISION_BITS = 32 - 8 - 2


def _bicubic(x: float) -> float:
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def resample_coeffs(in_size: int, out_size: int):
    """Pillow `precompute_coeffs` + `normalize_coeffs_8bpc` for the whole axis (box = full image): per output position the first
    input index, the tap count and the fixed-point taps."""
    support, scale = 2.0, in_size / out_size
    filterscale = max(scale, 1.0)
    support = support * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = [_bicubic((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            k = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + k * (1 << PRECISION_BITS)) if k < 0 else int(0.5 + k * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _pass(img: np.ndarray, out_size: int, axis: int) -> np.ndarray:
    """One separable pass over `axis` of a uint8 [H, W, C] image: sum of taps in int32 with the rounding constant, >> 22, clip."""
    bounds, kk = resample_coeffs(img.shape[axis], out_size)
    src = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.empty((out_size,) + src.shape[1:], dtype=np.uint8)
    for xx in range(out_size):
        x0, n = bounds[xx]
        acc = np.full(src.shape[1:], 1 << (PRECISION_BITS - 1), dtype=np.int64)
        for t in range(n):
            acc += src[x0 + t] * int(kk[xx, t])
        out[xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def pil_resize_bicubic(img: np.ndarray, out_w: int, out_h: int) -> np.ndarray:
    """uint8 [H, W, 3] -> uint8 [out_h, out_w, 3], Pillow `Image.resize((out_w, out_h), Image.BICUBIC)`: horizontal pass first (only
    when the width changes), then the vertical pass, each rounded to 8 bits."""
    h, w = img.shape[:2]
    if out_w != w:
        img = _pass(img, out_w, axis=1)
    if out_h != h:
        img = _pass(img, out_h, axis=0)
    return img


def resize_output_size(h: int, w: int, shortest_edge: int):
    """transformers 4.31 `get_resize_output_image_size(size=shortest_edge, default_to_square=False)`."""
    short, long = (w, h) if w <= h else (h, w)
    new_short, new_long = shortest_edge, int(shortest_edge * long / short)
    return (new_long, new_short) if w <= h else (new_short, new_long)  # (out_h, out_w)


def expand2square(img: np.ndarray, background) -> np.ndarray:
    """reference mm_utils.py:152-164 on a uint8 [H, W, 3] array."""
    h, w = img.shape[:2]
    if h == w:
        return img
    side = max(h, w)
    out = np.empty((side, side, 3), dtype=np.uint8)
    out[:] = np.asarray(background, dtype=np.uint8)
    top, left = (side - h) // 2, (side - w) // 2
    out[top:top + h, left:left + w] = img
    return out


def clip_preprocess(img: np.ndarray, size: int = 336, pad_to_square: bool = False, mean=CLIP_MEAN, std=CLIP_STD) -> np.ndarray:
    """uint8 RGB [H, W, 3] -> float32 [3, size, size]: the CLIPImageProcessor sequence of the reference's drivers."""
    if pad_to_square:
        img = expand2square(img, tuple(int(x * 255) for x in mean))
    h, w = img.shape[:2]
    out_h, out_w = resize_output_size(h, w, size)
    img = pil_resize_bicubic(img, out_w, out_h)
    top, left = (out_h - size) // 2, (out_w - size) // 2
    img = img[top:top + size, left:left + size]
    x = (img * (1 / 255)).astype(np.float32)                   # rescale: uint8 * python float -> float64 -> float32
    x = (x - np.asarray(mean, dtype=np.float32)) / np.asarray(std, dtype=np.float32)
    return np.ascontiguousarray(x.transpose(2, 0, 1))
