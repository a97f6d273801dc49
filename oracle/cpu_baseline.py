"""CPU timing of the reference's VDD path (TEST/BENCH INFRASTRUCTURE ONLY -- see oracle/vdd_oracle.py header).

The reference cannot be imported on the GPU box (needs /root/reference and transformers-4.31 names), so the timed
thing is the oracle port: literal, un-fused, one HF forward per stream per token (reference vcd_utils/vcd_sample.py:
109,163,178), bf16 weights on all host threads.  A full LLaVA-1.5-7B pass is minutes of CPU work per token batch, so
the sample is bounded: the real widths with a REDUCED layer count at two depths, extrapolated linearly in depth to the
real depth (every layer costs the same), for each phase (vision / prefill / per-token decode).
"""
from __future__ import annotations

import dataclasses
import os
import time

import torch


def _fast_state_dict(cfg, seed=0):
    """bf16 random weights straight from torch (numpy generation of 1e9 values is too slow for a bench leg)."""
    from llava_align_b200.weights import state_dict_shapes, _std_for
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape in state_dict_shapes(cfg).items():
        if name.endswith("weight") and len(shape) == 1:
            sd[name] = torch.ones(shape, dtype=torch.bfloat16)
        else:
            sd[name] = torch.empty(shape, dtype=torch.bfloat16).normal_(0.0, float(_std_for(name, shape, cfg, -1.0)), generator=g)
    return sd


def time_reference_path(cfg, B=1, T=32, n_new=3, use_dd=True, use_dd_unk=True, depths=(2, 4), vdepths=(2, 4), threads=None,
                        cd_alpha=1.0, cd_beta=0.1):
    """Returns dict(tokens_per_s, vision_s, prefill_s, decode_s_per_token, cores, sample) for the FULL-depth model,
    extrapolated from layer-reduced runs of the oracle."""
    from oracle.vdd_oracle import OracleVDD
    import numpy as np
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    rng = np.random.default_rng(0)
    ids = torch.from_numpy(rng.integers(3, cfg.vocab_size, size=(B, T)).astype("int64"))
    ids[:, 1 + T // 8] = -200
    img = torch.randn(B, 3, cfg.image_size, cfg.image_size).to(torch.bfloat16)
    meas = []
    for L, VL in zip(depths, vdepths):
        c = dataclasses.replace(cfg, num_hidden_layers=L, vision_num_hidden_layers=VL + 1)  # select_layer=-2 uses VL layers
        o = OracleVDD(c, _fast_state_dict(c), dtype=torch.bfloat16)
        with torch.no_grad():
            o.generate(ids, img, use_dd=use_dd, use_dd_unk=use_dd_unk, cd_alpha=cd_alpha, cd_beta=cd_beta, max_new_tokens=1,
                       eos_token_id=None)  # warm-up (thread pools, oneDNN primitives)
            t0 = time.perf_counter()
            out = o.generate(ids, img, use_dd=use_dd, use_dd_unk=use_dd_unk, cd_alpha=cd_alpha, cd_beta=cd_beta,
                             max_new_tokens=n_new, eos_token_id=None)
            wall = time.perf_counter() - t0
        t = out.timings
        meas.append((L, VL + 1, t["vision_s"], t["prefill_s"], t["decode_s"] / max(1, n_new - 1), wall))
        del o
    (L0, V0, v0, p0, d0, _), (L1, V1, v1, p1, d1, _) = meas
    lin = lambda x0, y0, x1, y1, x: y0 + (y1 - y0) * (x - x0) / (x1 - x0)
    Lf, Vf = cfg.num_hidden_layers, cfg.vision_num_hidden_layers
    vision = lin(V0, v0, V1, v1, Vf)        # the reference runs ALL ViT layers (clip_encoder.py:48), one is discarded
    prefill = lin(L0, p0, L1, p1, Lf)
    decode = lin(L0, d0, L1, d1, Lf)
    return dict(vision_s=vision, prefill_s=prefill, decode_s_per_token=decode, cores=threads,
                measured=[dict(llm_layers=m[0], vit_layers=m[1], vision_s=m[2], prefill_s=m[3], decode_s_per_token=m[4],
                               wall_s=m[5]) for m in meas],
                sample=(f"oracle port (HF eager bf16, un-fused: one forward per stream), real widths, B={B} T={T}, "
                        f"{n_new} tokens at LLM depth {depths} / ViT depth {tuple(v + 1 for v in vdepths)}, each phase "
                        f"extrapolated linearly to depth {Lf}/{Vf}"))


def tokens_per_s(r, B, n_new):
    return B * n_new / (r["vision_s"] + r["prefill_s"] + (n_new - 1) * r["decode_s_per_token"])
