"""Pins the CPU oracle to outputs of the UNMODIFIED reference (tests/golden/ref_tiny.*, made by
tests/golden/make_golden.py from /root/reference).  CPU only."""
import numpy as np
import pytest
import torch

from helpers import load_golden, make_inputs, tiny_model

META, DATA = load_golden()


@pytest.mark.parametrize("case", META["cases"], ids=[c["name"] for c in META["cases"]])
def test_oracle_matches_reference(case):
    cfg, sd, oracle = tiny_model(META["weight_seed"])
    ids, img = make_inputs(cfg, case["B"], case["T"], case["input_seed"])
    f, s = case["flags"], case["sampler"]
    torch.manual_seed(case["sample_seed"])  # the reference draws from the global generator (torch.multinomial)
    img_cd = torch.from_numpy(DATA[case["name"] + ".images_cd"]) if f.get("use_cd") else None
    out = oracle.generate(ids, img, images_cd=img_cd, use_dd=f.get("use_dd", False), use_dd_unk=f.get("use_dd_unk", False),
                          cd_alpha=f.get("cd_alpha"), cd_beta=f.get("cd_beta"),
                          temperature=s.get("temperature"), top_k=s.get("top_k"), top_p=s.get("top_p"),
                          max_new_tokens=case["n_new"], output_scores=True, eos_token_id=case["eos"])
    ref_seq = DATA[case["name"] + ".sequences"]
    ref_scores = DATA[case["name"] + ".scores"]
    assert out.sequences.shape == ref_seq.shape
    # sequences start with the untouched input ids incl. the -200 sentinel (llava_sampling.py:111-114)
    assert np.array_equal(out.sequences[:, :case["T"]].numpy(), ids.numpy())
    assert np.array_equal(out.sequences.numpy(), ref_seq), (out.sequences[:, case["T"]:], ref_seq[:, case["T"]:])
    got = torch.stack(out.scores, 0).numpy()
    assert got.shape == ref_scores.shape
    assert np.array_equal(np.isinf(got), np.isinf(ref_scores))  # same plausibility / top-k / top-p sets
    fin = np.isfinite(ref_scores)
    # both sides are fp32 on the same weights through the same HF layer math -> agree to rounding
    np.testing.assert_allclose(got[fin], ref_scores[fin], rtol=1e-4, atol=1e-4)


ATTN_CASES = [c for c in META["cases"] if c["name"] + ".attn0_last" in DATA.files]


@pytest.mark.parametrize("case", ATTN_CASES, ids=[c["name"] for c in ATTN_CASES])
def test_oracle_first_step_attention_matches_reference(case):
    """attentions[0][-1] (llava_calibrate.py:180): step 0, last layer, main stream, as returned by the unmodified reference."""
    cfg, sd, oracle = tiny_model(META["weight_seed"])
    ids, img = make_inputs(cfg, case["B"], case["T"], case["input_seed"])
    ref = DATA[case["name"] + ".attn0_last"]
    got = oracle.first_step_attention(ids, img).numpy()
    assert got.shape == ref.shape == (case["B"], cfg.num_attention_heads, case["T"] + cfg.num_patches - 1, case["T"] + cfg.num_patches - 1)
    np.testing.assert_allclose(got, ref, rtol=1e-5, atol=1e-6)
    assert np.all(np.triu(got[0, 0], 1) == 0)  # causal: nothing above the diagonal


def _attn_golden():
    import json
    import os
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    return json.load(open(os.path.join(here, "ref_tiny_attn.json"))), np.load(os.path.join(here, "ref_tiny_attn.npz"))


@pytest.mark.parametrize("name", [c["name"] for c in _attn_golden()[0]["cases"]])
def test_oracle_all_attentions_match_reference(name):
    """attentions[step][layer] for EVERY step and layer (vcd_sample.py:241-246) as the unmodified reference returned them
    (tests/golden/make_attn_golden.py): the oracle's teacher-forced replay of the main stream reproduces them."""
    meta, data = _attn_golden()
    case = next(c for c in META["cases"] if c["name"] == name)
    cfg, sd, oracle = tiny_model(META["weight_seed"])
    ids, img = make_inputs(cfg, case["B"], case["T"], case["input_seed"])
    new_tokens = torch.from_numpy(DATA[name + ".sequences"])[:, case["T"]:]
    got = oracle.all_attentions(ids, new_tokens, img)
    assert len(got) == case["n_new"]
    len0 = case["T"] + cfg.num_patches - 1
    for s, g in enumerate(got):
        ref = data[f"{name}.s{s}"]
        assert g.shape == ref.shape == (cfg.num_hidden_layers, case["B"], cfg.num_attention_heads, len0 if s == 0 else 1, len0 + s)
        np.testing.assert_allclose(g.numpy(), ref, rtol=1e-5, atol=1e-6)


def test_multinomial_is_exponential_race():
    """Appendix A Q6: torch.multinomial(p,1) == argmax(p / Exp(1)) drawn from the same generator state.
    The CUDA sampler relies on it: it receives q from torch and returns argmax(p/q)."""
    g = torch.Generator().manual_seed(3)
    p = torch.softmax(torch.randn(4, 1000, generator=g), -1)
    torch.manual_seed(11)
    a = torch.multinomial(p, 1).squeeze(1)
    torch.manual_seed(11)
    q = torch.empty_like(p).exponential_(1)
    b = torch.argmax(p / q, dim=-1)
    assert torch.equal(a, b)
