"""Seeded random-init weights in the HF LLaVA-1.5 state-dict layout.

No checkpoint can be downloaded, so every model in this repo is random-init (BASELINE.json configs say so).
The same state dict feeds the CPU oracle (fp32 math on bf16-rounded values) and the CUDA engine (bf16),
which is what makes their outputs comparable.  Key names follow the checkpoint the reference loads with
`LlavaLlamaForCausalLM.from_pretrained` (reference experiments/llava/model/builder.py:102-103):
`model.layers.N.self_attn.q_proj.weight`, `model.mm_projector.{0,2}.*`,
`model.vision_tower.vision_tower.vision_model.*`, `lm_head.weight`.
"""
from __future__ import annotations

import numpy as np
import torch

from .config import VddConfig

VT = "model.vision_tower.vision_tower.vision_model."


def _bf16_round(a: np.ndarray) -> torch.Tensor:
    return torch.from_numpy(a).to(torch.bfloat16).to(torch.float32)


def state_dict_shapes(cfg: VddConfig) -> dict:
    H, I, V = cfg.hidden_size, cfg.intermediate_size, cfg.vocab_size
    VH, VI, P = cfg.vision_hidden_size, cfg.vision_intermediate_size, cfg.patch_size
    s = {"model.embed_tokens.weight": (V, H), "model.norm.weight": (H,), "lm_head.weight": (V, H)}
    for i in range(cfg.num_hidden_layers):
        p = f"model.layers.{i}."
        for n in "qkvo":
            s[p + f"self_attn.{n}_proj.weight"] = (H, H)
        s[p + "mlp.gate_proj.weight"] = (I, H)
        s[p + "mlp.up_proj.weight"] = (I, H)
        s[p + "mlp.down_proj.weight"] = (H, I)
        s[p + "input_layernorm.weight"] = (H,)
        s[p + "post_attention_layernorm.weight"] = (H,)
    s["model.mm_projector.0.weight"] = (H, VH)
    s["model.mm_projector.0.bias"] = (H,)
    s["model.mm_projector.2.weight"] = (H, H)
    s["model.mm_projector.2.bias"] = (H,)
    s[VT + "embeddings.class_embedding"] = (VH,)
    s[VT + "embeddings.patch_embedding.weight"] = (VH, 3, P, P)
    s[VT + "embeddings.position_embedding.weight"] = (cfg.num_patches + 1, VH)
    s[VT + "pre_layrnorm.weight"] = (VH,)
    s[VT + "pre_layrnorm.bias"] = (VH,)
    for i in range(cfg.vision_num_hidden_layers):
        p = VT + f"encoder.layers.{i}."
        for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
            s[p + f"self_attn.{n}.weight"] = (VH, VH)
            s[p + f"self_attn.{n}.bias"] = (VH,)
        s[p + "layer_norm1.weight"] = (VH,)
        s[p + "layer_norm1.bias"] = (VH,)
        s[p + "layer_norm2.weight"] = (VH,)
        s[p + "layer_norm2.bias"] = (VH,)
        s[p + "mlp.fc1.weight"] = (VI, VH)
        s[p + "mlp.fc1.bias"] = (VI,)
        s[p + "mlp.fc2.weight"] = (VH, VI)
        s[p + "mlp.fc2.bias"] = (VH,)
    s[VT + "post_layernorm.weight"] = (VH,)
    s[VT + "post_layernorm.bias"] = (VH,)
    return s


def _std_for(name: str, shape, cfg: VddConfig, std: float) -> float:
    """N(0, std^2) for matrices.  lm_head gets 2.5/sqrt(H) so random-init logits have a usable spread
    (std ~2.5) and greedy margins are far above bf16 noise."""
    if name == "lm_head.weight":
        return 2.5 / np.sqrt(cfg.hidden_size)
    if name.endswith("patch_embedding.weight"):
        return 1.0 / np.sqrt(3 * cfg.patch_size ** 2)
    if name.endswith("embed_tokens.weight") or "embedding" in name:
        return 1.0
    if len(shape) >= 2:
        return std if std > 0 else 1.0 / np.sqrt(shape[-1])
    return 0.1  # biases


def make_state_dict(cfg: VddConfig, seed: int = 0, std: float = -1.0) -> dict:
    """Deterministic (numpy PCG64) fp32 state dict whose values are exactly representable in bf16.

    std<=0 selects fan-in scaling (1/sqrt(K)) so activations keep O(1) scale through random layers;
    std>0 uses that fixed std for all Linear matrices (the survey's N(0,0.02^2)).  Norm weights are
    1 + small noise so a missing norm-weight multiply is caught by parity tests.
    """
    rng = np.random.default_rng(seed)
    out = {}
    for name, shape in state_dict_shapes(cfg).items():
        if len(shape) == 1 and name.endswith("weight"):
            a = 1.0 + 0.1 * rng.standard_normal(shape, dtype=np.float32)
        else:
            a = _std_for(name, shape, cfg, std) * rng.standard_normal(shape, dtype=np.float32)
        out[name] = _bf16_round(a.astype(np.float32))
    return out


def make_state_dict_device(cfg: VddConfig, device, seed: int = 0, std: float = 0.02) -> dict:
    """Fast on-device random init for benchmarks at full size (7B/13B): bf16 tensors made by torch's CUDA
    generator.  Not bit-reproducible on CPU; parity tests use make_state_dict instead."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    out = {}
    for name, shape in state_dict_shapes(cfg).items():
        if name.endswith("weight") and len(shape) == 1:
            t = torch.ones(shape, device=device, dtype=torch.bfloat16)
        else:
            s = _std_for(name, shape, cfg, std)
            t = torch.empty(shape, device=device, dtype=torch.bfloat16).normal_(0.0, float(s), generator=g)
        out[name] = t
    return out
