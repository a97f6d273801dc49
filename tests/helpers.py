"""Shared synthetic-input builders for tests (same recipe as tests/golden/make_golden.py::make_inputs)."""
import json
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def make_inputs(cfg, B, T, seed):
    rng = np.random.default_rng(seed)
    ids = rng.integers(3, cfg.vocab_size, size=(B, T)).astype(np.int64)
    ids[:, 1 + T // 8] = -200
    img = rng.standard_normal((B, 3, cfg.image_size, cfg.image_size), dtype=np.float32)
    img = torch.from_numpy(img).to(torch.bfloat16).float()
    return torch.from_numpy(ids), img


def load_golden():
    with open(os.path.join(GOLDEN_DIR, "ref_tiny.json")) as f:
        meta = json.load(f)
    data = np.load(os.path.join(GOLDEN_DIR, "ref_tiny.npz"))
    return meta, data


_CACHE = {}


def tiny_model(seed=0):
    """(cfg, state_dict, OracleVDD) for the golden tiny config, cached per process."""
    from llava_align_b200.config import VddConfig
    from llava_align_b200.weights import make_state_dict
    from oracle.vdd_oracle import OracleVDD
    if seed not in _CACHE:
        meta, _ = load_golden()
        cfg = VddConfig(**meta["config"])
        sd = make_state_dict(cfg, seed=seed)
        _CACHE[seed] = (cfg, sd, OracleVDD(cfg, sd))
    return _CACHE[seed]


def prewarp_scores(stream_logits_step, cd_alpha, cd_beta):
    """The combined scores of one step BEFORE the warpers (vcd_sample.py:185,188-194) from the oracle's per-stream logits
    {"m", "u", "d"}: what a greedy (top_k=1) decision is taken on.  The post-warper `scores` of a top_k=1 run hold a single finite
    entry, so a top-2 margin has to come from here."""
    from oracle.vdd_oracle import contrastive_combine
    l, u, d = stream_logits_step["m"], stream_logits_step.get("u"), stream_logits_step.get("d")
    if u is not None and d is not None:
        cd = (u + d) / 2
    else:
        cd = u if u is not None else d
    return contrastive_combine(l, cd, cd_alpha, cd_beta) if cd is not None else l


def greedy_divergence_is_near_tie(got_tokens, ref_tokens, stream_logits, cd_alpha, cd_beta, tol):
    """got_tokens / ref_tokens: [B, n] new tokens.  Rows are followed until their first differing step; there the token the
    engine chose must score within `tol` of the oracle's best on the oracle's own pre-warp scores.  Returns the list of
    (row, step, gap) divergences (empty when the sequences are identical)."""
    out = []
    for b in range(ref_tokens.shape[0]):
        diff = (got_tokens[b] != ref_tokens[b]).nonzero()
        if diff.numel() == 0:
            continue
        st = int(diff[0, 0])
        s = prewarp_scores(stream_logits[st], cd_alpha, cd_beta)[b]
        gap = (s.max() - s[got_tokens[b, st]]).item()
        assert gap <= tol, (b, st, gap, got_tokens[b].tolist(), ref_tokens[b].tolist())
        out.append((b, st, gap))
    return out
