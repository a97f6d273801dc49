"""CPU timing of the reference's VDD path (TEST/BENCH INFRASTRUCTURE ONLY -- see oracle/vdd_oracle.py header).

The reference cannot be imported on the GPU box (needs /root/reference and transformers-4.31 names), so the timed
thing is the oracle port: literal, un-fused, one HF forward per stream per token (reference vcd_utils/vcd_sample.py:
109,163,178), bf16 weights on all host threads.  A full LLaVA-1.5-7B pass is minutes of CPU work per token batch, so
the sample is bounded: the real widths with a REDUCED layer count (8 of 32 LLM layers = 3.5 GB of bf16 weights,
well beyond the host's last-level cache so the decode leg stays DRAM-bound like the full model), each phase (vision /
prefill / per-token decode) scaled to the real depth by the parameters it touches.
"""
from __future__ import annotations

import dataclasses
import os
import time

import torch


def _fast_state_dict(cfg, seed=0):
    """bf16 random weights straight from torch.  Values are drawn once per distinct (shape, role) and cloned for the
    other layers: identical values across layers change nothing about the time a forward takes, and generating 3e9
    normal variates on the host would cost more than the measurement itself."""
    import re
    from llava_align_b200.weights import state_dict_shapes, _std_for
    g = torch.Generator().manual_seed(seed)
    sd, proto = {}, {}
    for name, shape in state_dict_shapes(cfg).items():
        if name.endswith("weight") and len(shape) == 1:
            sd[name] = torch.ones(shape, dtype=torch.bfloat16)
            continue
        role = re.sub(r"layers\.\d+\.", "layers.N.", name)
        if role not in proto:
            proto[role] = torch.empty(shape, dtype=torch.bfloat16).normal_(0.0, float(_std_for(name, shape, cfg, -1.0)), generator=g)
            sd[name] = proto[role]
        else:
            sd[name] = proto[role].clone()
    return sd


class ReferenceSample:
    """The layer-reduced oracle model, built once; run() times one bounded generate and scales it to full depth by the
    ratio of parameters touched (every phase is linear in them: decode is weight-bandwidth-bound, prefill and ViT are
    FLOP-bound)."""

    def __init__(self, cfg, B=1, T=32, n_new=3, use_dd=True, use_dd_unk=True, depth=8, vdepth=4, threads=None, cd_alpha=1.0,
                 cd_beta=0.1):
        from oracle.vdd_oracle import OracleVDD
        import numpy as np
        self.cfg, self.B, self.T, self.n_new = cfg, B, T, n_new
        self.threads = threads or os.cpu_count() or 1
        torch.set_num_threads(self.threads)
        rng = np.random.default_rng(0)
        self.ids = torch.from_numpy(rng.integers(3, cfg.vocab_size, size=(B, T)).astype("int64"))
        self.ids[:, 1 + T // 8] = -200
        self.img = torch.randn(B, 3, cfg.image_size, cfg.image_size).to(torch.bfloat16)
        self.depth = min(depth, cfg.num_hidden_layers)
        self.vdepth = min(vdepth, cfg.vision_num_hidden_layers - 1)
        c = dataclasses.replace(cfg, num_hidden_layers=self.depth, vision_num_hidden_layers=self.vdepth + 1)  # select_layer=-2
        self.oracle = OracleVDD(c, _fast_state_dict(c), dtype=torch.bfloat16)
        self.kw = dict(use_dd=use_dd, use_dd_unk=use_dd_unk, cd_alpha=cd_alpha, cd_beta=cd_beta, eos_token_id=None)
        with torch.no_grad():
            self.oracle.generate(self.ids, self.img, max_new_tokens=2, **self.kw)  # warm-up (thread pools, oneDNN primitives)

    def run(self):
        cfg = self.cfg
        with torch.no_grad():
            t0 = time.perf_counter()
            out = self.oracle.generate(self.ids, self.img, max_new_tokens=self.n_new, **self.kw)
            wall = time.perf_counter() - t0
        t = out.timings
        H, I, V = cfg.hidden_size, cfg.intermediate_size, cfg.vocab_size
        p_layer, p_head = 4 * H * H + 3 * H * I, V * H
        s_llm = (cfg.num_hidden_layers * p_layer + p_head) / (self.depth * p_layer + p_head)
        s_vit = cfg.vision_num_hidden_layers / (self.vdepth + 1)  # the reference runs ALL ViT layers (clip_encoder.py:48)
        dec = t["decode_s"] / max(1, self.n_new - 1)
        return dict(vision_s=t["vision_s"] * s_vit, prefill_s=t["prefill_s"] * s_llm, decode_s_per_token=dec * s_llm,
                    cores=self.threads,
                    measured=dict(llm_layers=self.depth, vit_layers=self.vdepth + 1, vision_s=t["vision_s"], prefill_s=t["prefill_s"],
                                  decode_s_per_token=dec, wall_s=wall, scale_llm=s_llm, scale_vit=s_vit),
                    sample=(f"oracle port (HF eager bf16, un-fused: one forward per stream), real widths, B={self.B} T={self.T}, "
                            f"{self.n_new} tokens with {self.depth} of {cfg.num_hidden_layers} LLM layers and {self.vdepth + 1} of "
                            f"{cfg.vision_num_hidden_layers} ViT layers; each phase scaled by parameters touched "
                            f"(x{s_llm:.2f} LLM, x{s_vit:.2f} ViT)"))


def time_reference_path(cfg, **kw):
    return ReferenceSample(cfg, **kw).run()


def tokens_per_s(r, B, n_new):
    return B * n_new / (r["vision_s"] + r["prefill_s"] + (n_new - 1) * r["decode_s_per_token"])
