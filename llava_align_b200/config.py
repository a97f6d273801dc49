"""Model dimensions for the VDD hot path (LLaVA-1.5 = CLIP ViT-L/14-336 + mlp2x_gelu projector + LLaMA).

The reference never states these numbers: it loads them from the HF checkpoints named in its
launch scripts (reference experiments/scripts/pope/run_llava.sh:6, experiments/llava/model/builder.py:97-103).
SURVEY.md Appendix B lists the public values used here.
"""
from __future__ import annotations

from dataclasses import dataclass, asdict

IMAGE_TOKEN_INDEX = -200  # reference experiments/llava/constants.py:9


@dataclass(frozen=True)
class VddConfig:
    # language model (HF LlamaConfig names)
    hidden_size: int = 4096
    intermediate_size: int = 11008
    num_hidden_layers: int = 32
    num_attention_heads: int = 32
    vocab_size: int = 32000
    max_position_embeddings: int = 4096
    rms_norm_eps: float = 1e-5
    rope_theta: float = 10000.0
    # vision tower (HF CLIPVisionConfig names)
    vision_hidden_size: int = 1024
    vision_intermediate_size: int = 4096
    vision_num_hidden_layers: int = 24
    vision_num_attention_heads: int = 16
    image_size: int = 336
    patch_size: int = 14
    vision_layer_norm_eps: float = 1e-5
    mm_vision_select_layer: int = -2  # reference clip_encoder.py:30; llava-v1.5 config
    # tokens
    pad_token_id: int = 0
    bos_token_id: int = 1
    eos_token_id: int = 2
    name: str = "llava-v1.5-7b"

    @property
    def head_dim(self) -> int:
        return self.hidden_size // self.num_attention_heads

    @property
    def vision_head_dim(self) -> int:
        return self.vision_hidden_size // self.vision_num_attention_heads

    @property
    def num_patches(self) -> int:
        return (self.image_size // self.patch_size) ** 2

    @property
    def vision_layers_used(self) -> int:
        """hidden_states[select_layer] with select_layer=-2 is the input of the last encoder layer,
        so only L-1 layers ever contribute (reference clip_encoder.py:29-37)."""
        sl = self.mm_vision_select_layer
        return self.vision_num_hidden_layers + 1 + sl if sl < 0 else sl

    def to_dict(self):
        return asdict(self)


LLAVA_15_7B = VddConfig()
LLAVA_15_13B = VddConfig(hidden_size=5120, intermediate_size=13824, num_hidden_layers=40,
                         num_attention_heads=40, name="llava-v1.5-13b")


def tiny_config(**kw) -> VddConfig:
    """Small model with the real head dims (128 / 64) so every CUDA kernel specialisation is exercised."""
    base = dict(hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=2,
                vocab_size=512, max_position_embeddings=512,
                vision_hidden_size=128, vision_intermediate_size=256, vision_num_hidden_layers=3,
                vision_num_attention_heads=2, image_size=56, patch_size=14, name="tiny")
    base.update(kw)
    return VddConfig(**base)


def by_name(name: str) -> VddConfig:
    """'llava-v1.5-7b', 'llava-v1.5-13b', 'tiny'; a suffix '@N' keeps the real widths but only N decoder layers
    (layer-reduced real-width models for parity tests against the CPU oracle)."""
    if "@" in name:
        from dataclasses import replace
        base, n_layers = name.rsplit("@", 1)
        return replace(by_name(base), num_hidden_layers=int(n_layers), name=name)
    n = name.lower()
    if "13b" in n:
        return LLAVA_15_13B
    if "tiny" in n:
        return tiny_config()
    return LLAVA_15_7B
