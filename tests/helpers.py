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
