"""Reading an HF LLaVA-1.5 checkpoint directory the way the reference's loader does
(reference experiments/llava/model/builder.py:97-103 -> LlavaLlamaForCausalLM.from_pretrained, then
`vision_tower.load_model()` at builder.py:137-141 -> CLIPVisionModel.from_pretrained(mm_vision_tower),
multimodal_encoder/clip_encoder.py:22-27).

A released LLaVA-1.5 checkpoint is two things on disk:
  * the LLaVA directory: config.json (LlamaConfig fields + mm_vision_tower, mm_hidden_size, mm_vision_select_layer, ...),
    weight shards (`*.safetensors` or `pytorch_model-*.bin`) holding model.embed_tokens / model.layers.* / model.norm /
    lm_head / model.mm_projector.*  -- the CLIP tower is NOT in these shards;
  * the CLIP directory named by `mm_vision_tower` (openai/clip-vit-large-patch14-336): config.json with `vision_config`
    (or a bare CLIPVisionConfig) and shards whose keys start with `vision_model.`.
This module turns both into (VddConfig, iterator of (engine tensor name, tensor)); nothing here needs a GPU.
"""
from __future__ import annotations

import glob
import json
import os
from typing import Iterator, Optional, Tuple

import torch

from .config import VddConfig
from .weights import VT

# keys a checkpoint may carry that hold no parameters of the path
_IGNORED_SUFFIXES = ("rotary_emb.inv_freq", "position_ids")
_CLIP_IGNORED_PREFIXES = ("text_model.", "text_projection", "visual_projection", "logit_scale")


def _read_json(path):
    with open(path) as f:
        return json.load(f)


def find_vision_tower_dir(model_path: str, llava_cfg: dict, vision_tower_path: Optional[str] = None) -> Optional[str]:
    """Where the CLIP weights live: explicit argument, else `mm_vision_tower` when it is a local directory, else a
    `vision_tower/` sub-directory of the checkpoint.  None => the shards themselves must hold model.vision_tower.* keys."""
    for cand in (vision_tower_path, llava_cfg.get("mm_vision_tower"), os.path.join(model_path, "vision_tower")):
        if cand and os.path.isdir(cand):
            return cand
    return None


def config_from_hf(model_path: str, vision_tower_path: Optional[str] = None) -> VddConfig:
    """VddConfig from the checkpoint's own config.json files (the reference reads the same numbers through
    AutoConfig / CLIPVisionConfig)."""
    c = _read_json(os.path.join(model_path, "config.json"))
    kw = dict(hidden_size=c["hidden_size"], intermediate_size=c["intermediate_size"], num_hidden_layers=c["num_hidden_layers"],
              num_attention_heads=c["num_attention_heads"], vocab_size=c["vocab_size"],
              max_position_embeddings=c.get("max_position_embeddings", 4096), rms_norm_eps=c.get("rms_norm_eps", 1e-5),
              rope_theta=c.get("rope_theta", 10000.0), mm_vision_select_layer=c.get("mm_vision_select_layer", -2),
              pad_token_id=c.get("pad_token_id", 0) or 0, bos_token_id=c.get("bos_token_id", 1), eos_token_id=c.get("eos_token_id", 2),
              name=c.get("_name_or_path") or os.path.basename(os.path.normpath(model_path)))
    if c.get("num_key_value_heads", c["num_attention_heads"]) != c["num_attention_heads"]:
        raise NotImplementedError("grouped-query attention is outside LLaVA-1.5 (num_key_value_heads != num_attention_heads)")
    if c.get("mm_projector_type", "mlp2x_gelu") != "mlp2x_gelu":
        raise NotImplementedError(f"mm_projector_type {c.get('mm_projector_type')!r}: only mlp2x_gelu (LLaVA-1.5) is built")
    if c.get("mm_vision_select_feature", "patch") != "patch":
        raise NotImplementedError("only mm_vision_select_feature='patch' is built (clip_encoder.py:32-33)")
    vdir = find_vision_tower_dir(model_path, c, vision_tower_path)
    if vdir is not None:
        v = _read_json(os.path.join(vdir, "config.json"))
        v = v.get("vision_config", v)
        kw.update(vision_hidden_size=v["hidden_size"], vision_intermediate_size=v["intermediate_size"],
                  vision_num_hidden_layers=v["num_hidden_layers"], vision_num_attention_heads=v["num_attention_heads"],
                  image_size=v["image_size"], patch_size=v["patch_size"], vision_layer_norm_eps=v.get("layer_norm_eps", 1e-5))
    if kw["hidden_size"] // kw["num_attention_heads"] != 128:
        raise NotImplementedError("the attention kernels are specialised for head_dim 128 (LLaVA-1.5 7B / 13B)")
    return VddConfig(**kw)


def _iter_shards(directory: str) -> Iterator[Tuple[str, torch.Tensor]]:
    st = sorted(glob.glob(os.path.join(directory, "*.safetensors")))
    if st:
        from safetensors import safe_open
        for f in st:
            with safe_open(f, framework="pt", device="cpu") as sf:
                for k in sf.keys():
                    yield k, sf.get_tensor(k)
        return
    bins = sorted(glob.glob(os.path.join(directory, "pytorch_model*.bin")))
    if not bins:
        raise FileNotFoundError(f"no *.safetensors or pytorch_model*.bin under {directory}")
    for f in bins:
        sd = torch.load(f, map_location="cpu", weights_only=True)
        for k, v in sd.items():
            yield k, v
        del sd


def map_llava_key(k: str) -> Optional[str]:
    """Checkpoint key -> engine tensor name (None: not a parameter of the path)."""
    if k.endswith(_IGNORED_SUFFIXES):
        return None
    return k


def map_clip_key(k: str) -> Optional[str]:
    if k.endswith(_IGNORED_SUFFIXES) or k.startswith(_CLIP_IGNORED_PREFIXES):
        return None
    if k.startswith("vision_model."):
        return VT + k[len("vision_model."):]
    return None


def iter_checkpoint_tensors(model_path: str, vision_tower_path: Optional[str] = None) -> Iterator[Tuple[str, torch.Tensor]]:
    """(engine tensor name, CPU tensor in the checkpoint's dtype) for every parameter of the path."""
    cfg = _read_json(os.path.join(model_path, "config.json"))
    for k, t in _iter_shards(model_path):
        name = map_llava_key(k)
        if name is not None:
            yield name, t
    vdir = find_vision_tower_dir(model_path, cfg, vision_tower_path)
    if vdir is not None and os.path.abspath(vdir) != os.path.abspath(model_path):
        for k, t in _iter_shards(vdir):
            name = map_clip_key(k)
            if name is not None:
                yield name, t
