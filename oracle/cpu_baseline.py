"""CPU timing of the reference's VDD path (TEST/BENCH INFRASTRUCTURE ONLY -- see oracle/vdd_oracle.py header).

The reference cannot be imported on the GPU box (needs /root/reference and transformers-4.31 names), so the timed
thing is the oracle port: literal, un-fused, one HF forward per stream per token (reference vcd_utils/vcd_sample.py:
109,163,178), bf16 weights on all host threads.  The model is built at its FULL depth (all 32 LLM layers, all 24 ViT
layers the reference runs); what is bounded is the NUMBER OF TOKENS of one sample: the image pass, the three prefills
and a few decode steps are each measured in full, nothing is scaled by parameter counts.  The tokens/s of the whole
256-token workload follows from those measured phases (decode steps are identical work up to a context that grows by
one token) and is labelled `projected`; the seconds a sample really took are reported beside it.
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
        sd[name] = proto[role]  # shared storage here; load_state_dict copies it into each layer's own parameter
    return sd


class ReferenceSample:
    """The full-depth oracle model, built once; run() times one bounded generate (image pass + prefill of every stream +
    n_new - 1 decode steps) and reports the measured phases.  `depth` / `vdepth` reduce the layer count only when asked to
    (CPU-only unit tests); the bench legs leave them at None = full depth."""

    def __init__(self, cfg, B=1, T=32, n_new=4, use_dd=True, use_dd_unk=True, depth=None, vdepth=None, threads=None, cd_alpha=1.0,
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
        self.depth = cfg.num_hidden_layers if depth is None else min(depth, cfg.num_hidden_layers)
        self.vlayers = cfg.vision_num_hidden_layers if vdepth is None else min(vdepth + 1, cfg.vision_num_hidden_layers)
        c = dataclasses.replace(cfg, num_hidden_layers=self.depth, vision_num_hidden_layers=self.vlayers)
        t0 = time.perf_counter()
        self.oracle = OracleVDD(c, _fast_state_dict(c), dtype=torch.bfloat16)
        self.kw = dict(use_dd=use_dd, use_dd_unk=use_dd_unk, cd_alpha=cd_alpha, cd_beta=cd_beta, eos_token_id=None)
        with torch.no_grad():
            self.oracle.generate(self.ids, self.img, max_new_tokens=2, **self.kw)  # warm-up (thread pools, oneDNN primitives)
        self.build_s = time.perf_counter() - t0

    def run(self):
        cfg = self.cfg
        with torch.no_grad():
            t0 = time.perf_counter()
            out = self.oracle.generate(self.ids, self.img, max_new_tokens=self.n_new, **self.kw)
            wall = time.perf_counter() - t0
        t = out.timings
        full = self.depth == cfg.num_hidden_layers and self.vlayers == cfg.vision_num_hidden_layers
        # a layer-reduced model (unit tests only) is scaled to the real depth by the parameters each phase touches
        H, I, V = cfg.hidden_size, cfg.intermediate_size, cfg.vocab_size
        p_layer, p_head = 4 * H * H + 3 * H * I, V * H
        s_llm = (cfg.num_hidden_layers * p_layer + p_head) / (self.depth * p_layer + p_head)
        s_vit = cfg.vision_num_hidden_layers / self.vlayers
        dec = t["decode_s"] / max(1, self.n_new - 1)
        return dict(vision_s=t["vision_s"] * s_vit, prefill_s=t["prefill_s"] * s_llm, decode_s_per_token=dec * s_llm,
                    cores=self.threads, wall_s=wall, full_depth=full,
                    measured=dict(llm_layers=self.depth, vit_layers=self.vlayers, vision_s=t["vision_s"], prefill_s=t["prefill_s"],
                                  decode_s_per_token=dec, wall_s=wall, scale_llm=s_llm, scale_vit=s_vit, build_s=self.build_s),
                    sample=(f"oracle port (HF eager bf16, un-fused: one forward per stream per token) of {cfg.name} at "
                            + (f"FULL depth ({self.depth} LLM + {self.vlayers} ViT layers)" if full else
                               f"{self.depth} of {cfg.num_hidden_layers} LLM and {self.vlayers} of {cfg.vision_num_hidden_layers} ViT "
                               f"layers, phases scaled by parameters touched (x{s_llm:.2f} / x{s_vit:.2f})")
                            + f", B={self.B} T={self.T}: image pass + {'3' if self.kw['use_dd'] and self.kw['use_dd_unk'] else '2'} prefills + "
                              f"{self.n_new - 1} decode steps measured ({wall:.1f} s); tokens/s of the whole workload projected from these phases"))


def time_reference_path(cfg, **kw):
    return ReferenceSample(cfg, **kw).run()


def tokens_per_s(r, B, n_new):
    return B * n_new / (r["vision_s"] + r["prefill_s"] + (n_new - 1) * r["decode_s_per_token"])
