"""`llava.model`: the engine-backed model class under the reference's name (reference llava/model/__init__.py:1)."""
from llava_align_b200.model import LlavaLlamaForCausalLM  # noqa: F401
