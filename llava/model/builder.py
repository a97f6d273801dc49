"""`llava.model.builder.load_pretrained_model` (reference experiments/llava/model/builder.py:26,148) -> the libvdd engine."""
from llava_align_b200.builder import load_pretrained_model  # noqa: F401
