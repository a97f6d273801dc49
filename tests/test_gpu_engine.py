"""GPU parity of the whole VDD path (ViT -> projector -> splice -> fused prefill of all streams -> fused decode steps ->
combine/sample) against the CPU oracle on the tiny seeded model, and against the reference's own golden outputs."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from helpers import load_golden, make_inputs, tiny_model

META, DATA = load_golden()
LOGIT_TOL = 6e-2   # bf16 activations / fp32 accumulate vs the fp32 oracle, logits of std ~2.5 (|l| up to ~12)


@pytest.fixture(scope="module")
def engine_model():
    from llava_align_b200.model import LlavaLlamaForCausalLM, evolve_vcd_sampling
    cfg, sd, oracle = tiny_model(META["weight_seed"])
    evolve_vcd_sampling()
    m = LlavaLlamaForCausalLM(cfg, "cuda:0")
    m.load_state_dict(sd)
    return cfg, m, oracle


def test_image_features(engine_model):
    from oracle.vdd_oracle import encode_images
    cfg, m, oracle = engine_model
    _, img = make_inputs(cfg, 2, 12, 5)
    ref = encode_images(oracle.vision, oracle.projector, img, cfg.mm_vision_select_layer)
    got = m.engine.encode_images(img).cpu()
    assert got.shape == ref.shape
    err = (got - ref).abs().max().item()
    assert err <= 3e-2 * max(1.0, ref.abs().max().item()), err


@pytest.mark.parametrize("flags", [(True, True), (True, False), (False, True), (False, False)])
@pytest.mark.parametrize("B", [1, 2])
def test_stream_logits_teacher_forced(engine_model, flags, B):
    """Every stream's last-position logits at every step, with the oracle's tokens fed to both sides."""
    cfg, m, oracle = engine_model
    use_unk, use_none = flags
    if B > 1 and use_none:
        pytest.skip("the reference's use_dd stream flattens the batch (B=1 only, SURVEY Appendix A Q2); covered at B=1")
    ids, img = make_inputs(cfg, B, 14, 31 + B)
    n = 6
    ref = oracle.generate(ids, img, use_dd=use_none, use_dd_unk=use_unk, cd_alpha=1.0, cd_beta=0.1, top_k=1,
                          max_new_tokens=n, keep_stream_logits=True, eos_token_id=None)
    eng = m.engine
    eng.encode_images(img)
    lg = eng.prefill(ids, True, use_unk, use_none, n)
    for step in range(n):
        got = lg.cpu()
        names = ["m"] + (["u"] if use_unk else []) + (["d"] if use_none else [])
        for si, nm in enumerate(names):
            r = ref.stream_logits[step][nm]
            err = (got[si] - r).abs().max().item()
            assert err <= LOGIT_TOL, (step, nm, err)
        if step + 1 < n:
            lg = eng.decode_step(ref.sequences[:, ids.shape[1] + step])


@pytest.mark.parametrize("case", META["cases"], ids=[c["name"] for c in META["cases"]])
def test_generate_matches_reference_golden(engine_model, case):
    """model.generate(...) through the fused CUDA path vs outputs of the UNMODIFIED reference (tests/golden)."""
    cfg, m, oracle = engine_model
    ids, img = make_inputs(cfg, case["B"], case["T"], case["input_seed"])
    f, s = case["flags"], case["sampler"]
    greedy = s.get("top_k") == 1
    B, n, V = case["B"], case["n_new"], cfg.vocab_size
    ref_seq = torch.from_numpy(DATA[case["name"] + ".sequences"])
    ref_scores = torch.from_numpy(DATA[case["name"] + ".scores"])
    if case["B"] > 1 and f.get("use_dd"):
        pytest.skip("reference use_dd is B=1 only")
    from llava_align_b200 import engine as E
    params = E.make_sample_params(cd_alpha=0.5 if f.get("cd_alpha") is None else f["cd_alpha"],
                                  cd_beta=0.1 if f.get("cd_beta") is None else f["cd_beta"],
                                  temperature=s.get("temperature"), top_k=s.get("top_k"), top_p=s.get("top_p"),
                                  use_contrast=bool(f.get("use_dd") or f.get("use_dd_unk") or f.get("use_cd")), eos_token_id=case["eos"],
                                  pad_token_id=cfg.pad_token_id)
    q = None
    if not greedy:
        # the same Exp(1) variates torch.multinomial drew inside the reference run (global CPU generator, seed 7)
        torch.manual_seed(case["sample_seed"])
        q = torch.stack([torch.empty(B, V).exponential_(1) for _ in range(n)]).cuda()
    scores = torch.empty(n, B, V, device="cuda")
    img_cd = torch.from_numpy(DATA[case["name"] + ".images_cd"]) if f.get("use_cd") else None
    toks, steps = m.engine.generate(ids, img, bool(f.get("use_dd_unk")), bool(f.get("use_dd")), params, n, q, scores, img_cd)
    assert steps == ref_seq.shape[1] - case["T"]
    got_scores = scores[:steps].cpu()
    a = 1.0 + (0.5 if f.get("cd_alpha") is None else f["cd_alpha"])
    tol = LOGIT_TOL * (2 * a) / (s.get("temperature") or 1.0)
    new_ref = ref_seq[:, case["T"]:]
    toks = toks.cpu()
    compared = 0
    for st in range(steps):
        # While both sides have emitted the same tokens they are in the same state: the post-warper scores of this step
        # (vcd_sample.py:200) must agree -- same finite set up to tokens sitting within bf16 noise of the cutoff / nucleus
        # edge, same values within the logit tolerance -- and so must the token drawn with the SAME variates q.
        fin = torch.isfinite(ref_scores[st])
        gfin = torch.isfinite(got_scores[st])
        assert (gfin != fin).sum().item() <= 2 * B, (st, (gfin != fin).sum())
        both = fin & gfin
        err = (got_scores[st][both] - ref_scores[st][both]).abs().max().item()
        assert err <= tol, (st, err)
        compared += 1
        if torch.equal(toks[:, st], new_ref[:, st]):
            continue
        # a different token is only legitimate at a near-tie of the reference's own decision: greedy = top-2 margin of the
        # scores; sampled = top-2 margin of the exponential race  score - log q  (argmax(softmax(s)/q) == argmax(s - log q))
        if q is None:
            # top_k=1 leaves ONE finite post-warper score: the margin comes from the oracle's pre-warp combined scores
            # (the oracle is pinned to these very golden vectors by tests/test_oracle_golden.py)
            from helpers import prewarp_scores
            o = oracle.generate(ids, img, images_cd=img_cd, use_dd=bool(f.get("use_dd")), use_dd_unk=bool(f.get("use_dd_unk")),
                                cd_alpha=f.get("cd_alpha"), cd_beta=f.get("cd_beta"), top_k=1, max_new_tokens=st + 1,
                                keep_stream_logits=True, eos_token_id=case["eos"])
            race = prewarp_scores(o.stream_logits[st], 0.5 if f.get("cd_alpha") is None else f["cd_alpha"],
                                  0.1 if f.get("cd_beta") is None else f["cd_beta"])
        else:
            race = ref_scores[st] - torch.log(q[st].cpu())
        rows = (toks[:, st] != new_ref[:, st]).nonzero()[:, 0]
        gap = (race[rows].max(dim=-1).values - race[rows, toks[rows, st]]).max().item()
        assert gap <= tol, (st, gap, toks[:, st], new_ref[:, st])
        break  # the sequences have legitimately diverged: later steps are no longer comparable
    assert compared >= 1


def test_generate_surface(engine_model):
    """The driver-level call (llava_sampling.py:96-114): layout of the result, kwargs validation, greedy bypass."""
    cfg, m, oracle = engine_model
    ids, img = make_inputs(cfg, 1, 12, 77)
    out = m.generate(ids.cuda(), images=img.half().cuda(), images_cd=None, cd_alpha=1.0, cd_beta=0.1, use_dd=True,
                     use_dd_unk=True, do_sample=True, temperature=1.0, top_p=None, top_k=1, max_new_tokens=6,
                     use_cache=True, eos_token_id=None)
    assert out.shape == (1, 12 + 6) and out.is_cuda
    assert torch.equal(out[:, :12].cpu(), ids)          # untouched prompt incl. the -200 sentinel
    ref = oracle.generate(ids, img, use_dd=True, use_dd_unk=True, cd_alpha=1.0, cd_beta=0.1, top_k=1, max_new_tokens=6,
                          eos_token_id=None, output_scores=True)
    if not torch.equal(out.cpu(), ref.sequences):
        first = int((out.cpu() != ref.sequences).nonzero()[0, 1]) - 12
        top2 = torch.topk(ref.scores[first], 2).values
        assert (top2[0, 0] - top2[0, 1]).item() < 0.3
    # do_sample=False: HF greedy_search, VDD flags silently ignored
    g1 = m.generate(ids.cuda(), images=img.cuda(), use_dd=True, use_dd_unk=True, do_sample=False, max_new_tokens=4, eos_token_id=None)
    g2 = m.generate(ids.cuda(), images=img.cuda(), do_sample=False, max_new_tokens=4, eos_token_id=None)
    assert torch.equal(g1, g2)
    r = m.generate(ids.cuda(), images=img.cuda(), use_dd_unk=True, do_sample=True, top_k=None, top_p=0.9, temperature=0.8,
                   max_new_tokens=5, output_scores=True, return_dict_in_generate=True, eos_token_id=None)
    assert r["sequences"].shape == (1, 17) and len(r["scores"]) == 5 and r["scores"][0].shape == (1, cfg.vocab_size)
    with pytest.raises(ValueError):
        m.generate(ids.cuda(), images=img.cuda(), bogus_kwarg=1, max_new_tokens=2)
    # use_cd: images_cd alone switches the contrast on (vcd_sample.py:122,147); identical images => (1+a)l - a*l with the cutoff
    c1 = m.generate(ids.cuda(), images=img.cuda(), images_cd=(img + 0.5 * torch.randn_like(img)).cuda(), cd_alpha=1.0, cd_beta=0.1,
                    do_sample=True, top_k=1, max_new_tokens=5, eos_token_id=None)
    assert c1.shape == (1, 17) and torch.equal(c1[:, :12].cpu(), ids)


ATTN_CASES = [c for c in META["cases"] if c["name"] + ".attn0_last" in DATA.files]


@pytest.mark.parametrize("case", ATTN_CASES, ids=[c["name"] for c in ATTN_CASES])
def test_first_step_attention_matches_reference(engine_model, case):
    """generate(output_attentions=True, return_dict_in_generate=True)['attentions'][0][-1] -- what llava_calibrate.py:180 reads --
    against the map the UNMODIFIED reference returned (tests/golden).  bf16 q/k on our side: absolute tolerance on probabilities."""
    cfg, m, oracle = engine_model
    ids, img = make_inputs(cfg, case["B"], case["T"], case["input_seed"])
    f = case["flags"]
    if case["B"] > 1 and f.get("use_dd"):
        pytest.skip("reference use_dd is B=1 only")
    ref = torch.from_numpy(DATA[case["name"] + ".attn0_last"])
    img_cd = torch.from_numpy(DATA[case["name"] + ".images_cd"]).cuda() if f.get("use_cd") else None
    out = m.generate(ids.cuda(), images=img.cuda(), images_cd=img_cd, use_dd=f.get("use_dd"), use_dd_unk=f.get("use_dd_unk"),
                     cd_alpha=f.get("cd_alpha"), cd_beta=f.get("cd_beta"), do_sample=True, top_k=1, max_new_tokens=3,
                     output_attentions=True, output_scores=True, return_dict_in_generate=True, eos_token_id=None)
    att = out["attentions"]
    assert len(att) == 3 and len(att[0]) == cfg.num_hidden_layers
    a = att[0][-1].cpu()
    assert a.shape == ref.shape
    assert (a - ref).abs().max().item() < 2e-2, (a - ref).abs().max().item()
    assert torch.allclose(a.sum(-1), torch.ones_like(a.sum(-1)), atol=1e-4)
    assert torch.equal(torch.triu(a[0, 0], 1), torch.zeros_like(a[0, 0]))
    torch.mean(a, dim=1).squeeze()  # what the driver does with it (llava_calibrate.py:182)
    # the tiny case fits the capture budget: the whole tuple is there, in HF's shapes
    len0 = ref.shape[-1]
    for s_ in range(3):
        assert len(att[s_]) == cfg.num_hidden_layers
        for l_ in range(cfg.num_hidden_layers):
            assert att[s_][l_].shape == (case["B"], cfg.num_attention_heads, len0 if s_ == 0 else 1, len0 + s_)


def _attn_golden():
    import json
    import os
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    return json.load(open(os.path.join(here, "ref_tiny_attn.json"))), np.load(os.path.join(here, "ref_tiny_attn.npz"))


@pytest.mark.parametrize("name", [c["name"] for c in _attn_golden()[0]["cases"]])
def test_all_attentions_match_reference(engine_model, name):
    """attentions[step][layer] for EVERY step and layer of generate(output_attentions=True) (vcd_sample.py:241-246) against the
    tuples the UNMODIFIED reference returned (tests/golden/make_attn_golden.py), while the greedy tokens agree with it.  bf16
    keys and bf16 hidden states on our side: absolute tolerance on probabilities."""
    cfg, m, oracle = engine_model
    meta, data = _attn_golden()
    case = next(c for c in META["cases"] if c["name"] == name)
    ids, img = make_inputs(cfg, case["B"], case["T"], case["input_seed"])
    f = case["flags"]
    out = m.generate(ids.cuda(), images=img.cuda(), use_dd=f.get("use_dd"), use_dd_unk=f.get("use_dd_unk"), cd_alpha=f.get("cd_alpha"),
                     cd_beta=f.get("cd_beta"), do_sample=True, top_k=1, max_new_tokens=case["n_new"], output_attentions=True,
                     return_dict_in_generate=True, eos_token_id=None)
    att = out["attentions"]
    assert len(att) == case["n_new"]
    ref_new = torch.from_numpy(DATA[name + ".sequences"])[:, case["T"]:]
    got_new = out["sequences"][:, case["T"]:].cpu()
    checked = 0
    for s_ in range(case["n_new"]):
        # step s_ attends over the tokens generated before it: comparable while every row's earlier tokens agree
        if s_ > 0 and not torch.equal(got_new[:, :s_], ref_new[:, :s_]):
            break
        ref = torch.from_numpy(data[f"{name}.s{s_}"])
        got = torch.stack([a.cpu() for a in att[s_]], dim=0)
        assert got.shape == ref.shape
        assert (got - ref).abs().max().item() < 2e-2, (s_, (got - ref).abs().max().item())
        assert torch.allclose(got.sum(-1), torch.ones_like(got.sum(-1)), atol=1e-4)
        checked += 1
    assert checked >= 3, (checked, got_new, ref_new)  # a near-tie may end the comparison early, not before a few decode steps


def test_attentions_last_only_mode(engine_model, monkeypatch):
    """VDD_ATTENTIONS=last (what the automatic budget falls back to for long generations): only attentions[0][-1] exists."""
    cfg, m, oracle = engine_model
    ids, img = make_inputs(cfg, 1, 12, 102)
    monkeypatch.setenv("VDD_ATTENTIONS", "last")
    out = m.generate(ids.cuda(), images=img.cuda(), use_dd=True, use_dd_unk=True, do_sample=True, top_k=1, max_new_tokens=3,
                     output_attentions=True, return_dict_in_generate=True, eos_token_id=None)
    att = out["attentions"]
    assert att[0][-1].shape[-1] == 12 + cfg.num_patches - 1
    with pytest.raises(NotImplementedError):
        att[1][-1]
    with pytest.raises(NotImplementedError):
        att[0][0]
    monkeypatch.setenv("VDD_ATTENTIONS", "all")
    out2 = m.generate(ids.cuda(), images=img.cuda(), use_dd=True, use_dd_unk=True, do_sample=True, top_k=1, max_new_tokens=3,
                      output_attentions=True, return_dict_in_generate=True, eos_token_id=None)
    assert torch.equal(out2["sequences"], out["sequences"])  # the eager capture steps compute what the graph computes
    assert (out2["attentions"][0][-1] - att[0][-1]).abs().max().item() == 0.0


@pytest.mark.parametrize("B,use_unk,use_none,with_image,T", [(2, True, True, True, 10), (3, True, True, True, 9),
                                                            (4, True, False, True, 7), (1, False, False, False, 5),
                                                            (2, True, False, False, 6), (5, True, True, True, 4),
                                                            (50, True, True, True, 5)])
def test_row_shapes_against_oracle(engine_model, B, use_unk, use_none, with_image, T):
    """Row counts on both sides of the GEMV / tensor-core decode switch (R = 6, 9, 8, 1, 4, 15) and beyond one 128-row tile
    (R = 150: no split-K planes, two M tiles), text-only prompts, and the
    row-wise definition of the use_dd stream for B > 1, teacher-forced on the oracle's tokens."""
    cfg, m, oracle = engine_model
    ids, img = make_inputs(cfg, B, T, 200 + B + T)
    if not with_image:
        ids[ids < 0] = 7
    n = 4
    ref = oracle.generate(ids, img if with_image else None, use_dd=use_none, use_dd_unk=use_unk, cd_alpha=1.0, cd_beta=0.1, top_k=1,
                          max_new_tokens=n, keep_stream_logits=True, eos_token_id=None)
    eng = m.engine
    if with_image:
        eng.encode_images(img)
    lg = eng.prefill(ids, with_image, use_unk, use_none, n)
    names = ["m"] + (["u"] if use_unk else []) + (["d"] if use_none else [])
    for step in range(n):
        got = lg.cpu()
        for si, nm in enumerate(names):
            err = (got[si] - ref.stream_logits[step][nm]).abs().max().item()
            assert err <= LOGIT_TOL, (step, nm, err)
        if step + 1 < n:
            lg = eng.decode_step(ref.sequences[:, T + step])
    out = m.generate(ids, images=img if with_image else None, use_dd=use_none, use_dd_unk=use_unk, cd_alpha=1.0, cd_beta=0.1,
                     do_sample=True, top_k=1, max_new_tokens=1, eos_token_id=None)
    assert out.shape == (B, T + 1)


@pytest.mark.parametrize("B,use_unk,use_none", [(1, True, True), (1, False, True), (2, True, False), (1, False, False)])
def test_whole_step_kernel_against_oracle(monkeypatch, B, use_unk, use_none):
    """The opt-in whole-step persistent kernel (VDD_STEP=1, decode_step.cu: TMA-fed weight rings across all phases of a decode
    step, attention as a phase of the same launch) for 3 / 2 / 4 / 1 fused rows: teacher-forced stream logits against the
    oracle, and its greedy VDD tokens against the oracle's (exact, or a near-tie of the oracle's own pre-warp scores)."""
    from llava_align_b200.model import LlavaLlamaForCausalLM, evolve_vcd_sampling
    cfg, sd, oracle = tiny_model(META["weight_seed"])
    evolve_vcd_sampling()
    monkeypatch.setenv("VDD_STEP", "1")  # read at vdd_create
    m = LlavaLlamaForCausalLM(cfg, "cuda:0")
    monkeypatch.delenv("VDD_STEP")
    m.load_state_dict(sd)
    T, n = 11, 6
    ids, img = make_inputs(cfg, B, T, 400 + B)
    ref = oracle.generate(ids, img, use_dd=use_none, use_dd_unk=use_unk, cd_alpha=1.0, cd_beta=0.1, top_k=1, max_new_tokens=n,
                          keep_stream_logits=True, eos_token_id=None)
    eng = m.engine
    eng.encode_images(img)
    lg = eng.prefill(ids, True, use_unk, use_none, n)
    names = ["m"] + (["u"] if use_unk else []) + (["d"] if use_none else [])
    for step in range(n):
        got = lg.cpu()
        for si, nm in enumerate(names):
            err = (got[si] - ref.stream_logits[step][nm]).abs().max().item()
            assert err <= LOGIT_TOL, (step, nm, err)
        if step + 1 < n:
            lg = eng.decode_step(ref.sequences[:, T + step])
    kw = dict(images=img, use_dd=use_none, use_dd_unk=use_unk, cd_alpha=1.0, cd_beta=0.1, do_sample=True, top_k=1, max_new_tokens=n,
              eos_token_id=None)
    out = m.generate(ids, **kw)
    assert eng.launch_count(reset=True) > 0
    from helpers import greedy_divergence_is_near_tie
    greedy_divergence_is_near_tie(out[:, T:].cpu(), ref.sequences[:, T:], ref.stream_logits, 1.0, 0.1, 0.3)


def test_stream_scores_fold_the_extra_generates_of_the_calibrate_driver(engine_model):
    """llava_calibrate.py:192-194 runs two more generates per question on the image-free prompts (sentinel deleted / replaced by
    <unk>) only to read `scores[0]`.  generate(..., return_stream_scores=True) returns those scores from the fused batch: compare
    with the oracle's separate text-only generates on the very prompts the debias streams stand for."""
    cfg, m, oracle = engine_model
    from oracle.vdd_oracle import debias_ids
    ids, img = make_inputs(cfg, 2, 13, 91)
    kw = dict(cd_alpha=1.0, cd_beta=0.1, do_sample=True, temperature=0.7, top_k=None, top_p=None, max_new_tokens=3, eos_token_id=None)
    out = m.generate(ids.cuda(), images=img.cuda(), use_dd=True, use_dd_unk=True, output_scores=True, return_dict_in_generate=True,
                     return_stream_scores=True, **kw)
    ss = out["stream_scores"]
    assert set(ss) == {"unk", "none"} and ss["unk"].shape == (2, cfg.vocab_size)
    for name in ("unk", "none"):
        ref = oracle.generate(debias_ids(ids, name), None, temperature=0.7, max_new_tokens=1, output_scores=True, eos_token_id=None)
        got = ss[name].cpu()
        assert torch.isfinite(got).all()
        assert (got - ref.scores[0]).abs().max().item() <= LOGIT_TOL / 0.7
    # and what the driver does with them (utils/metrics.py:102-113): softmax -> top-10
    probs = torch.softmax(ss["none"], dim=-1).float().cpu()
    assert torch.topk(probs, k=10).indices.shape == (2, 10)
    # without the flag nothing extra is returned; with no debias stream there is nothing to return
    assert "stream_scores" not in m.generate(ids.cuda(), images=img.cuda(), use_dd=True, return_dict_in_generate=True, **kw)


def test_ragged_batch_rows_equal_the_same_prompts_alone(engine_model):
    """Ragged batches: left- and right-padded rows under an attention_mask compute exactly what the unpadded
    prompts compute alone -- every stream's logits at every step against the oracle run row by row on the unpadded prompts,
    and generate() end to end.  (The reference itself cannot batch ragged prompts: llava_arch.py:91-94.)"""
    cfg, m, oracle = engine_model
    T, n = 14, 4
    ids, img = make_inputs(cfg, 3, T, 61)          # sentinel at column 1 + T // 8 = 2
    mask = torch.ones(3, T, dtype=torch.long)
    ids[1, :3] = cfg.pad_token_id; mask[1, :3] = 0    # left padding ...
    ids[1, 5] = -200; ids[1, 2] = cfg.pad_token_id    # ... with the sentinel moved inside the real part
    ids[2, 9:] = cfg.pad_token_id; mask[2, 9:] = 0    # right padding
    rows = [ids[b][mask[b] != 0].unsqueeze(0) for b in range(3)]
    refs = [oracle.generate(rows[b], img[b:b + 1], use_dd=True, use_dd_unk=True, cd_alpha=1.0, cd_beta=0.1, top_k=1, max_new_tokens=n,
                            keep_stream_logits=True, eos_token_id=None) for b in range(3)]
    eng = m.engine
    eng.encode_images(img)
    lg = eng.prefill(ids, True, True, True, n, attention_mask=mask)
    for step in range(n):
        got = lg.cpu()
        for b in range(3):
            for si, nm in enumerate("mud"):
                err = (got[si, b] - refs[b].stream_logits[step][nm][0]).abs().max().item()
                assert err <= LOGIT_TOL, (step, b, nm, err)
        if step + 1 < n:
            lg = eng.decode_step(torch.stack([refs[b].sequences[0, rows[b].shape[1] + step] for b in range(3)]))
    out = m.generate(ids.cuda(), images=img.cuda(), attention_mask=mask.cuda(), use_dd=True, use_dd_unk=True, cd_alpha=1.0, cd_beta=0.1,
                     do_sample=True, top_k=1, max_new_tokens=n, eos_token_id=None)
    assert out.shape == (3, T + n) and torch.equal(out[:, :T].cpu(), ids)
    from helpers import greedy_divergence_is_near_tie
    for b in range(3):
        ref_new = refs[b].sequences[:, rows[b].shape[1]:]
        greedy_divergence_is_near_tie(out[b:b + 1, T:].cpu(), ref_new, refs[b].stream_logits, 1.0, 0.1, 0.3)
    with pytest.raises(ValueError):   # a masked-out sentinel is a caller error
        bad = mask.clone(); bad[0, 2] = 0
        m.generate(ids.cuda(), images=img.cuda(), attention_mask=bad.cuda(), use_dd=True, do_sample=True, max_new_tokens=2, eos_token_id=None)


@pytest.mark.parametrize("B,sampler", [(1, dict(top_k=None, top_p=None, temperature=1.0)), (2, dict(top_k=None, top_p=0.9, temperature=0.8)),
                                       (2, dict(top_k=1))])
def test_generate_draws_the_variates_torch_multinomial_would(engine_model, B, sampler):
    """model.generate() takes its randomness from torch's CUDA generator like the reference (torch.multinomial per step): the
    tokens equal a run that is handed the per-step `empty(B, V).exponential_(1)` draws explicitly, and the generator ends where
    that loop leaves it."""
    cfg, m, oracle = engine_model
    from llava_align_b200 import engine as E
    ids, img = make_inputs(cfg, B, 12, 700 + B)
    n, V = 7, cfg.vocab_size
    gen = torch.cuda.default_generators[torch.cuda.current_device()]
    torch.manual_seed(99)
    qs = torch.stack([torch.empty(B, V, device="cuda").exponential_(1) for _ in range(n)])
    end_offset = gen.get_offset()
    params = E.make_sample_params(cd_alpha=1.0, cd_beta=0.1, temperature=sampler.get("temperature"), top_k=sampler.get("top_k"),
                                  top_p=sampler.get("top_p"), use_contrast=True, eos_token_id=None, pad_token_id=cfg.pad_token_id)
    want, steps = m.engine.generate(ids, img, True, True, params, n, qs, None)
    torch.manual_seed(99)
    out = m.generate(ids.cuda(), images=img.cuda(), use_dd=True, use_dd_unk=True, cd_alpha=1.0, cd_beta=0.1, do_sample=True,
                     max_new_tokens=n, eos_token_id=None, **sampler)
    assert torch.equal(out[:, 12:].cpu(), want.cpu())
    assert gen.get_offset() == end_offset
