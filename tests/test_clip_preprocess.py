"""Image preprocessing row: the CPU oracle against Pillow itself and against the committed golden vectors
(tests/golden/make_clip_preprocess_golden.py: Pillow + the transformers-4.31 CLIPImageProcessor sequence + the reference's own
expand2square), and -- on the GPU -- libvdd's kernels against the oracle, bit for bit."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

from oracle import clip_preprocess_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
META = json.load(open(os.path.join(GOLD, "clip_preprocess.json")))
DATA = np.load(os.path.join(GOLD, "clip_preprocess.npz"))

def _image(name):
    """Stored for the small cases; the 336-pixel cases are regenerated from their (H, W, seed) with the generator's recipe."""
    if name + ".image" in DATA.files:
        return DATA[name + ".image"]
    m = META[name]
    rng = np.random.default_rng(m["seed"])
    smooth = rng.random((m["H"] // 8 + 2, m["W"] // 8 + 2, 3))
    img = np.kron(smooth, np.ones((8, 8, 1)))[:m["H"], :m["W"]] * 200 + rng.random((m["H"], m["W"], 3)) * 55
    return img.astype(np.uint8)


SIZES = [(40, 64, 21, 34), (37, 53, 80, 111), (100, 100, 336, 336), (640, 427, 503, 336), (64, 48, 64, 30), (33, 77, 20, 77),
         (5, 9, 14, 3)]


@pytest.mark.parametrize("h,w,oh,ow", SIZES)
def test_oracle_resize_is_pillow_bit_exact(h, w, oh, ow):
    from PIL import Image
    a = (np.random.default_rng(h * 1000 + w).random((h, w, 3)) * 255).astype(np.uint8)
    ref = np.asarray(Image.fromarray(a).resize((ow, oh), Image.BICUBIC))
    assert np.array_equal(O.pil_resize_bicubic(a, ow, oh), ref)


@pytest.mark.parametrize("name", sorted(META))
def test_oracle_matches_golden(name):
    m = META[name]
    out = O.clip_preprocess(_image(name), m["size"], pad_to_square=m["pad"])
    assert out.shape == (3, m["size"], m["size"]) and out.dtype == np.float32
    assert hashlib.sha256(out.tobytes()).hexdigest() == m["sha256"]
    if name + ".pixel_values" in DATA.files:
        assert np.array_equal(out, DATA[name + ".pixel_values"])


def test_oracle_expand2square_matches_reference_semantics():
    a = (np.random.default_rng(3).random((30, 50, 3)) * 255).astype(np.uint8)
    bg = tuple(int(x * 255) for x in O.CLIP_MEAN)
    sq = O.expand2square(a, bg)
    assert sq.shape == (50, 50, 3) and np.array_equal(sq[10:40], a) and (sq[:10] == np.array(bg, dtype=np.uint8)).all()
    assert np.array_equal(O.expand2square(a.transpose(1, 0, 2), bg)[:, 10:40], a.transpose(1, 0, 2))
    assert O.resize_output_size(480, 640, 336) == (336, 448) and O.resize_output_size(640, 427, 336) == (503, 336)


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(META))
def test_gpu_preprocess_matches_golden_bit_exact(name):
    from llava_align_b200.hostside import GpuClipImageProcessor
    m = META[name]
    proc = GpuClipImageProcessor(m["size"])
    img = _image(name)
    for src in (torch.from_numpy(img), torch.from_numpy(img).cuda()):  # host pixels (copied inside) and device-resident pixels
        out = proc.preprocess(src, return_tensors="pt", pad_to_square=m["pad"])["pixel_values"][0]
        assert out.is_cuda and out.dtype == torch.float32 and out.shape == (3, m["size"], m["size"])
        assert hashlib.sha256(out.cpu().numpy().tobytes()).hexdigest() == m["sha256"], name


@pytest.mark.gpu
@pytest.mark.parametrize("h,w,size,pad", [(31, 45, 56, False), (500, 333, 336, False), (333, 500, 336, True), (336, 336, 336, False),
                                          (1200, 1600, 336, False), (17, 900, 56, True), (224, 224, 336, False)])
def test_gpu_preprocess_matches_oracle_on_random_sizes(h, w, size, pad):
    from PIL import Image
    from llava_align_b200.hostside import GpuClipImageProcessor, process_images
    a = (np.random.default_rng(h + w).random((h, w, 3)) * 255).astype(np.uint8)
    ref = O.clip_preprocess(a, size, pad_to_square=pad)
    proc = GpuClipImageProcessor(size)
    pil = Image.fromarray(a)
    if pad:  # the drivers' process_images(..., image_aspect_ratio='pad') route (mm_utils.py:166-173)
        out = process_images([pil], proc, type("Cfg", (), {"image_aspect_ratio": "pad"})())[0]
    else:    # the drivers' direct call (llava_sampling.py:82)
        out = proc.preprocess(pil, return_tensors="pt")["pixel_values"][0]
    assert np.array_equal(out.cpu().numpy(), ref)
    # what the drivers do next: .unsqueeze(0).half().cuda()
    assert out.unsqueeze(0).half().cuda().shape == (1, 3, size, size)
