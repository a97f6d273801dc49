"""Drop-in for the reference's `vcd_utils/vcd_sample.py`: the drivers do

    from vcd_utils.vcd_sample import evolve_vcd_sampling
    evolve_vcd_sampling()

at import time (reference experiments/eval/sampling/llava_sampling.py:24-25).  In the reference this monkey-patches
transformers' GenerationMixin.sample with the VDD loop (vcd_sample.py:325-326); here it arms the same loop inside
libvdd for models created by llava_align_b200 (the loop itself is CUDA: llava_align_b200/csrc/engine.cu).
"""
from llava_align_b200.model import evolve_vcd_sampling  # noqa: F401
