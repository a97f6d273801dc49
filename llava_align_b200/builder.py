"""`load_pretrained_model(model_path, model_base, model_name) -> (tokenizer, model, image_processor, context_len)`
with the reference's signature and return tuple (reference experiments/llava/model/builder.py:26,148).

No checkpoint can be fetched offline, so two sources are supported:
  * "random:<name>[:seed]"  -- seeded random-init weights of the named architecture (llava-v1.5-7b / -13b / tiny);
  * a local directory holding an HF LLaVA-1.5 checkpoint (config.json + *.safetensors or pytorch_model*.bin shards) whose
    CLIP tower is found through `mm_vision_tower` / `vision_tower_path` (llava_align_b200/checkpoint.py).
LoRA / 8-bit / 4-bit / MPT branches of the reference loader are out of scope.
"""
from __future__ import annotations

import os

import torch

from .config import by_name
from .model import LlavaLlamaForCausalLM
from .weights import state_dict_shapes, make_state_dict, _std_for


def _load_random(model: LlavaLlamaForCausalLM, seed: int, exact: bool):
    cfg = model.vdd_config
    if exact:  # bit-reproducible on CPU (parity tests)
        model.engine.load_state_dict(make_state_dict(cfg, seed))
        return
    g = torch.Generator(device=model.device)
    g.manual_seed(seed)
    for name, shape in state_dict_shapes(cfg).items():
        if name.endswith("weight") and len(shape) == 1:
            t = torch.ones(shape, device=model.device, dtype=torch.bfloat16)
        else:
            t = torch.empty(shape, device=model.device, dtype=torch.float32).normal_(0.0, float(_std_for(name, shape, cfg, -1.0)), generator=g)
        model.engine.load_tensor(name, t)
        del t
    model.engine.check_complete()


def load_pretrained_model(model_path, model_base=None, model_name=None, load_8bit=False, load_4bit=False, device_map="auto",
                          device="cuda", tp_size=1, tp_rank=0, vision_tower_path=None):
    if load_8bit or load_4bit:
        raise NotImplementedError("bitsandbytes loading is outside the VDD path")
    dev = torch.device("cuda", torch.cuda.current_device()) if device == "cuda" else torch.device(device)
    tokenizer = image_processor = None
    if str(model_path).startswith("random:"):
        parts = str(model_path).split(":")
        cfg = by_name(parts[1])
        seed = int(parts[2]) if len(parts) > 2 else 0
        model = LlavaLlamaForCausalLM(cfg, dev, tp_size, tp_rank)
        _load_random(model, seed, exact=cfg.name == "tiny")
        from .hostside import GpuClipImageProcessor
        image_processor = GpuClipImageProcessor(cfg.image_size, device=dev)  # no HF files to read a processor config from
        model.get_vision_tower().image_processor = image_processor
    else:
        # local checkpoint directory (reference builder.py:97-103,137-141): LLaVA shards + the CLIP tower they point at
        from .checkpoint import config_from_hf, iter_checkpoint_tensors
        cfg = config_from_hf(model_path, vision_tower_path)
        model = LlavaLlamaForCausalLM(cfg, dev, tp_size, tp_rank)
        for name, t in iter_checkpoint_tensors(model_path, vision_tower_path):
            model.engine.load_tensor(name, t)
        model.engine.check_complete()
        import logging
        from .checkpoint import find_vision_tower_dir, _read_json
        hf_cfg = _read_json(os.path.join(model_path, "config.json"))
        # reference builder.py:143-146: context_len = config.max_sequence_length if the checkpoint defines it, else 2048
        # (released LLaVA-1.5 configs do not define it)
        if "max_sequence_length" in hf_cfg:
            model.config.max_sequence_length = int(hf_cfg["max_sequence_length"])
        # from_pretrained also reads generation_config.json into model.generation_config (defaults of generate())
        gpath = os.path.join(model_path, "generation_config.json")
        if os.path.exists(gpath):
            for k, v in _read_json(gpath).items():
                if hasattr(model.generation_config, k):
                    setattr(model.generation_config, k, v)
        try:
            from transformers import AutoTokenizer, CLIPImageProcessor
            tokenizer = AutoTokenizer.from_pretrained(model_path, use_fast=False)
            vdir = find_vision_tower_dir(model_path, hf_cfg, vision_tower_path)
            image_processor = CLIPImageProcessor.from_pretrained(vdir or model_path)
            model.get_vision_tower().image_processor = image_processor
        except (FileNotFoundError, OSError, ValueError) as ex:
            # weights-only directories (tests) carry no tokenizer / preprocessor files; anything else propagates
            logging.getLogger(__name__).warning("no tokenizer / image processor under %s (%s): returning None for them",
                                                model_path, type(ex).__name__)
    context_len = getattr(model.config, "max_sequence_length", 2048)  # reference builder.py:143-146
    return tokenizer, model, image_processor, context_len
