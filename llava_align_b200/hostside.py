"""Host-side helpers of the VDD callers: what the reference's drivers do around `model.generate` on the
CPU -- prompt templating, `<image>` tokenisation, model naming -- restated so the drivers run against this engine on a
box that has no copy of the reference tree.  When the reference tree IS available the `llava/` shim package serves the
reference's own `llava.conversation` / `llava.mm_utils` untouched instead (see llava/__init__.py).

Reference behaviour followed here:
  constants                      experiments/llava/constants.py:7-12
  tokenizer_image_token          experiments/llava/mm_utils.py:185-204
  get_model_name_from_path       experiments/llava/mm_utils.py:207-213
  KeywordsStoppingCriteria       experiments/llava/mm_utils.py:215-247 (drivers build it, never pass it: llava_sampling.py:91-109)
  expand2square / process_images experiments/llava/mm_utils.py:152-182
  disable_torch_init             experiments/llava/utils.py:93-99
  llava_v1 / vicuna_v1 prompts   experiments/llava/conversation.py:29-60,252-262,335-345
"""
from __future__ import annotations

import dataclasses
import enum
import sys
import types
from typing import List, Optional, Sequence

import torch

IGNORE_INDEX = -100
IMAGE_TOKEN_INDEX = -200
DEFAULT_IMAGE_TOKEN = "<image>"
DEFAULT_IMAGE_PATCH_TOKEN = "<im_patch>"
DEFAULT_IM_START_TOKEN = "<im_start>"
DEFAULT_IM_END_TOKEN = "<im_end>"


# ------------------------------------------------------------------------------------------------ prompt -> ids
def tokenizer_image_token(prompt: str, tokenizer, image_token_index: int = IMAGE_TOKEN_INDEX, return_tensors: Optional[str] = None):
    """Tokenise the text around every '<image>' marker separately and join the pieces with ONE sentinel id per marker.
    Every piece comes back from the tokenizer with its own BOS; only the first is kept."""
    pieces = [list(tokenizer(part).input_ids) for part in prompt.split(DEFAULT_IMAGE_TOKEN)]
    bos = getattr(tokenizer, "bos_token_id", None)
    has_bos = bool(pieces) and bool(pieces[0]) and pieces[0][0] == bos
    skip = 1 if has_bos else 0
    ids: List[int] = [pieces[0][0]] if has_bos else []
    for i, piece in enumerate(pieces):
        if i:
            ids.append(image_token_index)
        ids.extend(piece[skip:])
    if return_tensors is None:
        return ids
    if return_tensors != "pt":
        raise ValueError(f"Unsupported tensor type: {return_tensors}")
    return torch.tensor(ids, dtype=torch.long)


def get_model_name_from_path(model_path: str) -> str:
    parts = model_path.strip("/").split("/")
    return parts[-2] + "_" + parts[-1] if parts[-1].startswith("checkpoint-") else parts[-1]


def disable_torch_init():
    """The reference skips torch's default Linear/LayerNorm initialisation before from_pretrained; the engine never builds
    torch modules, so there is nothing to skip."""
    return None


class KeywordsStoppingCriteria:
    """Constructed by every driver and then NOT passed to generate (llava_sampling.py:91-109): kept for the constructor
    contract (keyword ids without a leading BOS, start length) and usable as a predicate on a finished sequence."""

    def __init__(self, keywords: Sequence[str], tokenizer, input_ids: torch.Tensor):
        self.keywords = list(keywords)
        self.tokenizer = tokenizer
        self.start_len = input_ids.shape[1]
        self.keyword_ids = []
        for kw in self.keywords:
            ids = list(tokenizer(kw).input_ids)
            if len(ids) > 1 and ids[0] == getattr(tokenizer, "bos_token_id", None):
                ids = ids[1:]
            self.keyword_ids.append(torch.tensor(ids))
        self.max_keyword_len = max((len(k) for k in self.keyword_ids), default=0)

    def __call__(self, output_ids: torch.Tensor, scores=None, **kwargs) -> bool:
        done = []
        for row in output_ids:
            hit = any(len(k) <= row.shape[0] and torch.equal(row[-len(k):].cpu(), k) for k in self.keyword_ids if len(k))
            if not hit:
                tail = row[-min(row.shape[0] - self.start_len, self.max_keyword_len):] if row.shape[0] > self.start_len else row[:0]
                text = self.tokenizer.batch_decode(tail[None], skip_special_tokens=True)[0]
                hit = any(kw in text for kw in self.keywords)
            done.append(hit)
        return all(done)


# ------------------------------------------------------------------------------------------------ images
def expand2square(pil_img, background_color):
    """Centre the image on a square canvas of the longer side (LLaVA-1.5 `image_aspect_ratio == 'pad'`)."""
    from PIL import Image
    w, h = pil_img.size
    if w == h:
        return pil_img
    side = max(w, h)
    canvas = Image.new(pil_img.mode, (side, side), background_color)
    canvas.paste(pil_img, ((side - w) // 2, (side - h) // 2))
    return canvas


CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


class GpuClipImageProcessor:
    """`image_processor` of the drivers, on the GPU: `preprocess(image, return_tensors='pt')['pixel_values'][0]` (reference
    llava_sampling.py:82, llava_calibrate.py:149) returns the same float32 values as the transformers-4.31 CLIPImageProcessor the
    reference loads (builder.py:137-141) -- Pillow BICUBIC shortest-edge resize, center crop, 1/255, CLIP mean / std -- computed
    by libvdd (csrc/preprocess.cu) and left on the device (the drivers' `.half().cuda()` is then a device-side cast).
    There is no CPU path: without the CUDA library this raises."""

    def __init__(self, size=336, image_mean=CLIP_MEAN, image_std=CLIP_STD, device=None):
        self.size = {"shortest_edge": int(size)}
        self.crop_size = {"height": int(size), "width": int(size)}
        self.image_mean, self.image_std = list(image_mean), list(image_std)
        self.device = device

    def _one(self, image, pad):
        import ctypes as C
        import numpy as np
        from . import _lib
        lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.VddError("GpuClipImageProcessor needs a CUDA device (sm_100a); there is no CPU path")
        dev = torch.device(self.device) if self.device is not None else torch.device("cuda", torch.cuda.current_device())
        if isinstance(image, torch.Tensor):
            arr = image
        else:
            if hasattr(image, "convert"):
                image = image.convert("RGB")  # do_convert_rgb
            arr = torch.from_numpy(np.ascontiguousarray(np.asarray(image)))
        if arr.dtype != torch.uint8 or arr.dim() != 3 or arr.shape[2] != 3:
            raise ValueError(f"expected a uint8 RGB image [H, W, 3], got {tuple(arr.shape)} {arr.dtype}")
        arr = arr.contiguous()
        H, W = int(arr.shape[0]), int(arr.shape[1])
        S = self.size["shortest_edge"]
        out = torch.empty(3, S, S, dtype=torch.float32, device=dev)
        mean = (C.c_float * 3)(*self.image_mean)
        std = (C.c_float * 3)(*self.image_std)
        bg = (C.c_uint8 * 3)(*[int(x * 255) for x in self.image_mean]) if pad else None  # mm_utils.py:171
        with torch.cuda.device(dev):
            _lib.check(lib.vdd_clip_preprocess(C.c_void_p(arr.data_ptr()), H, W, int(arr.is_cuda), S, bg, mean, std,
                                               C.c_void_p(out.data_ptr()), 0, C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)),
                       "vdd_clip_preprocess")
            if not arr.is_cuda:
                torch.cuda.current_stream(dev).synchronize()  # the host pixels are read asynchronously
        return out

    def preprocess(self, images, return_tensors="pt", pad_to_square=False, **kwargs):
        many = isinstance(images, (list, tuple))
        outs = [self._one(im, pad_to_square) for im in (images if many else [images])]
        return {"pixel_values": torch.stack(outs, dim=0)}

    __call__ = preprocess


def process_images(images, image_processor, model_cfg):
    mode = getattr(model_cfg, "image_aspect_ratio", None)
    if mode == "anyres":
        raise NotImplementedError("anyres multi-crop inputs are LLaVA-1.6; the VDD path is one 336px image per prompt")
    if mode != "pad":
        return image_processor(images, return_tensors="pt")["pixel_values"]
    if isinstance(image_processor, GpuClipImageProcessor):  # the padding is folded into the resize kernel's source fetch
        return image_processor.preprocess(list(images), return_tensors="pt", pad_to_square=True)["pixel_values"]
    bg = tuple(int(x * 255) for x in image_processor.image_mean)
    out = [image_processor.preprocess(expand2square(im, bg), return_tensors="pt")["pixel_values"][0] for im in images]
    return torch.stack(out, dim=0) if all(o.shape == out[0].shape for o in out) else out


# ------------------------------------------------------------------------------------------------ prompt templates
class SeparatorStyle(enum.Enum):
    SINGLE = enum.auto()
    TWO = enum.auto()
    MPT = enum.auto()
    PLAIN = enum.auto()
    LLAMA_2 = enum.auto()


@dataclasses.dataclass
class Conversation:
    system: str
    roles: Sequence[str]
    messages: List[List[str]]
    offset: int = 0
    sep_style: SeparatorStyle = SeparatorStyle.SINGLE
    sep: str = "###"
    sep2: Optional[str] = None
    version: str = "Unknown"

    def append_message(self, role, message):
        self.messages.append([role, message])

    def copy(self):
        return dataclasses.replace(self, messages=[[r, m] for r, m in self.messages])

    def get_prompt(self) -> str:
        if self.sep_style == SeparatorStyle.TWO:
            seps = (self.sep, self.sep2)
        elif self.sep_style == SeparatorStyle.SINGLE:
            seps = (self.sep, self.sep)
        else:
            raise NotImplementedError(f"{self.sep_style}: only the SINGLE / TWO styles of the LLaVA-1.5 drivers are restated here")
        out = self.system + seps[0]
        for i, (role, msg) in enumerate(self.messages):
            out += f"{role}: {msg}{seps[i % 2]}" if msg else f"{role}:"
        return out


_V1_SYSTEM = ("A chat between a curious human and an artificial intelligence assistant. "
              "The assistant gives helpful, detailed, and polite answers to the human's questions.")
_VICUNA_SYSTEM = ("A chat between a curious user and an artificial intelligence assistant. "
                  "The assistant gives helpful, detailed, and polite answers to the user's questions.")
conv_templates = {
    "llava_v1": Conversation(system=_V1_SYSTEM, roles=("USER", "ASSISTANT"), messages=[], sep_style=SeparatorStyle.TWO, sep=" ",
                             sep2="</s>", version="v1"),
    "vicuna_v1": Conversation(system=_VICUNA_SYSTEM, roles=("USER", "ASSISTANT"), messages=[], sep_style=SeparatorStyle.TWO,
                              sep=" ", sep2="</s>", version="v1"),
}
conv_templates["v1"] = conv_templates["vicuna_v1"]


# ------------------------------------------------------------------------------------------------ module registration
def register_fallback_modules(pkg: str = "llava"):
    """Serve `llava.constants`, `llava.mm_utils`, `llava.utils`, `llava.conversation` from this file (no reference tree)."""
    me = sys.modules[__name__]
    groups = {
        "constants": ("IGNORE_INDEX", "IMAGE_TOKEN_INDEX", "DEFAULT_IMAGE_TOKEN", "DEFAULT_IMAGE_PATCH_TOKEN",
                      "DEFAULT_IM_START_TOKEN", "DEFAULT_IM_END_TOKEN"),
        "mm_utils": ("tokenizer_image_token", "get_model_name_from_path", "KeywordsStoppingCriteria", "expand2square",
                     "process_images", "GpuClipImageProcessor"),
        "utils": ("disable_torch_init",),
        "conversation": ("conv_templates", "SeparatorStyle", "Conversation"),
    }
    for sub, names in groups.items():
        full = f"{pkg}.{sub}"
        if full in sys.modules:
            continue
        mod = types.ModuleType(full)
        mod.__doc__ = f"{full}: served by llava_align_b200.hostside (no reference tree on this box)"
        for n in names:
            setattr(mod, n, getattr(me, n))
        sys.modules[full] = mod
        if pkg in sys.modules:
            setattr(sys.modules[pkg], sub, mod)
