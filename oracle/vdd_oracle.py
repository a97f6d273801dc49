"""CPU oracle for the Visual-Debias-Decoding (VDD) hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module; the product path (llava_align_b200/) never does and fails loudly without its CUDA library.

It restates, literally and un-fused (one forward per stream, exactly like the reference), what
/root/reference does on this path:

  * the sampling step            reference vcd_utils/vcd_sample.py:93-299
  * the debias-input construction reference vcd_utils/vcd_sample.py:147-161,171-177
  * the contrastive combine       reference vcd_utils/vcd_sample.py:185,188-194
  * warp + softmax + multinomial  reference vcd_utils/vcd_sample.py:197-202
  * the image splice              reference experiments/llava/model/llava_arch.py:122-163,194-202
  * the vision tower              reference experiments/llava/model/multimodal_encoder/clip_encoder.py:29-51
  * the projector                 reference experiments/llava/model/multimodal_projector/builder.py:39-46

The layer arithmetic itself is NOT in /root/reference: it lives in the un-vendored `transformers`
dependency (unpinned in the reference's requirements.txt:2; its `sample(...)` signature pins the 4.31-4.33
API).  The oracle uses the installed transformers (5.5.0) LlamaForCausalLM (eager attention) and
CLIPVisionModel as that arithmetic, in fp32, on bf16-representable weights.

Parity pin: tests/golden/make_golden.py imports the UNMODIFIED reference modules
(vcd_utils/vcd_sample.py, experiments/llava/model/{llava_arch,language_model/llava_llama}.py,
multimodal_encoder/clip_encoder.py, multimodal_projector/builder.py) in the authoring container with a few
name shims for the newer transformers, runs them on tiny seeded models and commits the outputs under
tests/golden/.  tests/test_oracle_golden.py checks this oracle against those vectors, so the oracle is
pinned to outputs of the reference itself (the reference ships no tests or golden vectors of its own).
"""
from __future__ import annotations

import math
import time
from dataclasses import dataclass, field
from typing import List, Optional

import torch
from torch import nn

IMAGE_TOKEN_INDEX = -200  # reference experiments/llava/constants.py:9


# --------------------------------------------------------------------------------------------------
# model construction (HF modules as layer arithmetic)
# --------------------------------------------------------------------------------------------------
def build_modules(cfg, state_dict, dtype=torch.float32):
    """Returns (llama: LlamaForCausalLM, vision: CLIPVisionModel, projector: nn.Sequential)."""
    from transformers import LlamaConfig, LlamaForCausalLM, CLIPVisionConfig, CLIPVisionModel

    lc = LlamaConfig(hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size,
                     num_hidden_layers=cfg.num_hidden_layers, num_attention_heads=cfg.num_attention_heads,
                     num_key_value_heads=cfg.num_attention_heads, vocab_size=cfg.vocab_size,
                     max_position_embeddings=cfg.max_position_embeddings, rms_norm_eps=cfg.rms_norm_eps,
                     rope_theta=cfg.rope_theta, hidden_act="silu", attention_bias=False, mlp_bias=False,
                     tie_word_embeddings=False, pad_token_id=cfg.pad_token_id, bos_token_id=cfg.bos_token_id,
                     eos_token_id=cfg.eos_token_id, attn_implementation="eager")
    vc = CLIPVisionConfig(hidden_size=cfg.vision_hidden_size, intermediate_size=cfg.vision_intermediate_size,
                          num_hidden_layers=cfg.vision_num_hidden_layers,
                          num_attention_heads=cfg.vision_num_attention_heads, image_size=cfg.image_size,
                          patch_size=cfg.patch_size, hidden_act="quick_gelu",
                          layer_norm_eps=cfg.vision_layer_norm_eps, attn_implementation="eager")
    with torch.device("meta"):
        llama = LlamaForCausalLM(lc)
        vision = CLIPVisionModel(vc)
    llama = llama.to_empty(device="cpu").to(dtype)
    vision = vision.to_empty(device="cpu").to(dtype)
    # reference multimodal_projector/builder.py:39-46  (mlp2x_gelu)
    projector = nn.Sequential(nn.Linear(cfg.vision_hidden_size, cfg.hidden_size), nn.GELU(),
                              nn.Linear(cfg.hidden_size, cfg.hidden_size)).to(dtype)

    vt_prefix = "model.vision_tower.vision_tower."
    pj_prefix = "model.mm_projector."
    sd_l, sd_v, sd_p = {}, {}, {}
    for k, v in state_dict.items():
        v = v.detach().to("cpu", dtype)
        if k.startswith(vt_prefix):
            sd_v[k[len(vt_prefix):]] = v
        elif k.startswith(pj_prefix):
            sd_p[k[len(pj_prefix):]] = v
        else:
            sd_l[k] = v
    missing = llama.load_state_dict(sd_l, strict=False)
    assert not [m for m in missing.missing_keys if "rotary" not in m], missing
    _reinit_rotary(llama)
    mv = vision.load_state_dict(sd_v, strict=False)
    assert not [m for m in mv.missing_keys if "position_ids" not in m], mv
    vision.vision_model.embeddings.position_ids = torch.arange(cfg.num_patches + 1).unsqueeze(0)
    projector.load_state_dict(sd_p)
    return llama.eval(), vision.eval(), projector.eval()


def _reinit_rotary(llama):
    """to_empty() leaves the (non-persistent) rotary inv_freq buffer uninitialised; rebuild the module."""
    rot = llama.model.rotary_emb
    llama.model.rotary_emb = type(rot)(config=llama.config)


# --------------------------------------------------------------------------------------------------
# vision + splice
# --------------------------------------------------------------------------------------------------
@torch.no_grad()
def encode_images(vision, projector, images, select_layer=-2):
    """reference llava_arch.py:82-85 -> clip_encoder.py:39-51: hidden_states[select_layer][:, 1:] -> projector."""
    out = vision(images.to(next(vision.parameters()).dtype), output_hidden_states=True)
    feats = out.hidden_states[select_layer][:, 1:]
    return projector(feats)


@torch.no_grad()
def splice_main(embed_tokens, input_ids, image_features):
    """reference llava_arch.py:122-163,194-195 for the one-sentinel-per-row, equal-length case the drivers
    produce: replace the single -200 position by the row's image feature rows -> [B, T-1+P, H]."""
    rows = []
    for b in range(input_ids.shape[0]):
        ids = input_ids[b]
        pos = torch.where(ids == IMAGE_TOKEN_INDEX)[0]
        if pos.numel() == 0:  # llava_arch.py:126-137: row without an image keeps its text embeddings
            rows.append(embed_tokens(ids))
            continue
        p = int(pos[0])
        rows.append(torch.cat([embed_tokens(ids[:p]), image_features[b], embed_tokens(ids[p + 1:])], dim=0))
    return torch.stack(rows, dim=0)


def debias_ids(input_ids, kind):
    """reference vcd_sample.py:153-160.  kind 'unk': sentinel -> token 0; kind 'none': sentinel deleted
    (the reference flattens the batch here, so it is defined for B=1; row-wise deletion equals it at B=1)."""
    if kind == "unk":
        ids = input_ids.clone()
        ids[ids == IMAGE_TOKEN_INDEX] = 0
        return ids
    keep = [row[row != IMAGE_TOKEN_INDEX] for row in input_ids]
    return torch.stack(keep, dim=0)


# --------------------------------------------------------------------------------------------------
# the step: combine + warp + sample   (fp32; the reference runs it in fp16)
# --------------------------------------------------------------------------------------------------
def contrastive_combine(logits, logits_cd, cd_alpha, cd_beta):
    """reference vcd_sample.py:191-194."""
    cutoff = torch.log(torch.tensor(cd_beta)) + logits.max(dim=-1, keepdim=True).values
    diffs = (1 + cd_alpha) * logits - cd_alpha * logits_cd
    return diffs.masked_fill(logits < cutoff, -float("inf"))


def warp(scores, temperature=None, top_k=None, top_p=None):
    """HF 4.31 `_get_logits_warper` order: Temperature (if !=1) -> TopK (if set, !=0) -> TopP (if <1);
    semantics of transformers/generation/logits_process.py Temperature/TopK/TopP warpers, restated."""
    if temperature is not None and temperature != 1.0:
        scores = scores / temperature
    if top_k is not None and top_k != 0:
        k = min(int(top_k), scores.shape[-1])
        kth = torch.topk(scores, k)[0][..., -1, None]
        scores = scores.masked_fill(scores < kth, -float("inf"))
    if top_p is not None and top_p < 1.0:
        sorted_logits, sorted_indices = torch.sort(scores, descending=False, stable=True)
        cumulative_probs = sorted_logits.softmax(dim=-1).cumsum(dim=-1)
        sorted_remove = cumulative_probs <= (1 - top_p)
        sorted_remove[..., -1:] = False  # min_tokens_to_keep = 1
        remove = sorted_remove.scatter(-1, sorted_indices, sorted_remove)
        scores = scores.masked_fill(remove, -float("inf"))
    return scores


def sample_from_scores(scores, q=None):
    """reference vcd_sample.py:201-202: softmax then torch.multinomial(probs, 1).
    ATen draws q ~ Exp(1) per element and returns argmax(probs / q); passing the same q reproduces it."""
    probs = torch.softmax(scores, dim=-1)
    if q is None:
        return torch.multinomial(probs, num_samples=1).squeeze(1), probs
    return torch.argmax(probs / q, dim=-1), probs


def vdd_step(l, u=None, d=None, cd_alpha=0.5, cd_beta=0.1, temperature=None, top_k=None, top_p=None, q=None):
    """One row-batch of the step: l main logits, u unk-stream logits, d none-stream logits (each [B,V] or None).
    Returns (tokens [B], scores [B,V])."""
    if u is not None and d is not None:
        cd = (u + d) / 2  # vcd_sample.py:185
    else:
        cd = u if u is not None else d
    s = contrastive_combine(l, cd, cd_alpha, cd_beta) if cd is not None else l
    s = warp(s, temperature, top_k, top_p)
    tok, _ = sample_from_scores(s, q)
    return tok, s


# --------------------------------------------------------------------------------------------------
# the generation loop, literal: separate forwards and separate KV caches per stream
# --------------------------------------------------------------------------------------------------
@dataclass
class OracleOutput:
    sequences: torch.Tensor
    scores: Optional[List[torch.Tensor]] = None
    stream_logits: Optional[List[dict]] = None  # per step {"m":..., "u":..., "d":...}
    timings: dict = field(default_factory=dict)


class OracleVDD:
    """LLaVA-1.5 + the reference sampler, on CPU."""

    def __init__(self, cfg, state_dict, dtype=torch.float32):
        self.cfg = cfg
        self.dtype = dtype
        self.llama, self.vision, self.projector = build_modules(cfg, state_dict, dtype)

    @torch.no_grad()
    def _forward(self, cache, input_ids=None, inputs_embeds=None):
        from transformers.cache_utils import DynamicCache
        if cache is None:
            cache = DynamicCache()
        out = self.llama(input_ids=input_ids, inputs_embeds=inputs_embeds, past_key_values=cache, use_cache=True)
        return out.logits[:, -1, :].float(), cache  # vcd_sample.py:119 (last position only)

    @torch.no_grad()
    def first_step_attention(self, input_ids, images=None):
        """`attentions[0][-1]` of the reference's generate(output_attentions=True) (vcd_sample.py:241-246, read at
        llava_calibrate.py:180): post-softmax weights of the LAST layer for the main stream's prefill, [B, heads, len, len].
        Runs on a deep copy of the module: under transformers 5.5 a forward that records attentions leaves state behind
        that perturbs later forwards of the same module (see tests/golden/make_golden.py)."""
        import copy
        llama = copy.deepcopy(self.llama)
        embed = llama.model.embed_tokens
        if images is not None:
            feats = encode_images(self.vision, self.projector, images, self.cfg.mm_vision_select_layer)
            out = llama(inputs_embeds=splice_main(embed, input_ids, feats.to(self.dtype)), use_cache=True, output_attentions=True)
        else:
            out = llama(input_ids=input_ids, use_cache=True, output_attentions=True)
        return out.attentions[-1].float()

    @torch.no_grad()
    def all_attentions(self, input_ids, new_tokens, images=None):
        """`attentions[step]` for every step of the reference's generate(output_attentions=True) (vcd_sample.py:241-246 appends the
        MAIN stream's `outputs.attentions` each step): step 0 is the prefill, [L, B, heads, len, len]; step s >= 1 is the forward
        of generated token s-1 over the main KV cache, [L, B, heads, 1, len + s].  `new_tokens` [B, n] are the tokens the run
        produced (the maps do not depend on how they were sampled).  Deep copy for the same reason as first_step_attention."""
        import copy
        from transformers.cache_utils import DynamicCache
        llama = copy.deepcopy(self.llama)
        cache = DynamicCache()
        if images is not None:
            feats = encode_images(self.vision, self.projector, images, self.cfg.mm_vision_select_layer)
            out = llama(inputs_embeds=splice_main(llama.model.embed_tokens, input_ids, feats.to(self.dtype)), past_key_values=cache,
                        use_cache=True, output_attentions=True)
        else:
            out = llama(input_ids=input_ids, past_key_values=cache, use_cache=True, output_attentions=True)
        steps = [torch.stack([a.float() for a in out.attentions], dim=0)]
        for s in range(1, new_tokens.shape[1]):
            out = llama(input_ids=new_tokens[:, s - 1:s], past_key_values=cache, use_cache=True, output_attentions=True)
            steps.append(torch.stack([a.float() for a in out.attentions], dim=0))
        return steps

    @torch.no_grad()
    def generate(self, input_ids, images=None, images_cd=None, use_dd=False, use_dd_unk=False, cd_alpha=None, cd_beta=None,
                 temperature=None, top_k=None, top_p=None, max_new_tokens=16, do_sample=True,
                 output_scores=False, keep_stream_logits=False, q_list=None, eos_token_id="cfg",
                 pad_token_id=None):
        cfg = self.cfg
        eos = cfg.eos_token_id if eos_token_id == "cfg" else eos_token_id
        pad = cfg.pad_token_id if pad_token_id is None else pad_token_id
        cd_alpha = 0.5 if cd_alpha is None else cd_alpha  # vcd_sample.py:188
        cd_beta = 0.1 if cd_beta is None else cd_beta     # vcd_sample.py:189
        B = input_ids.shape[0]
        t = {}
        t0 = time.perf_counter()
        embed = self.llama.model.embed_tokens
        if images is not None:
            feats = encode_images(self.vision, self.projector, images, cfg.mm_vision_select_layer)
            main_embeds = splice_main(embed, input_ids, feats.to(self.dtype))
        else:
            main_embeds = None
        cd_embeds = None
        if images_cd is not None and images is not None:
            # use_cd (vcd_sample.py:122,148-150): the contrast forward sees the same prompt with images_cd spliced in
            # With use_dd / use_dd_unk also set, the `if use_cd:` arm at vcd_sample.py:148 still wins the first contrast
            # forward (the unk stream is never run); only when BOTH flags are set does :171-185 add the 'none' stream.
            feats_cd = encode_images(self.vision, self.projector, images_cd, cfg.mm_vision_select_layer)
            cd_embeds = splice_main(embed, input_ids, feats_cd.to(self.dtype))
        t["vision_s"] = time.perf_counter() - t0
        vdd = bool(use_dd or use_dd_unk or cd_embeds is not None) and do_sample  # do_sample=False bypasses the patched sampler (Q9)
        unfinished = torch.ones(B, dtype=torch.long)
        caches = {"m": None, "u": None, "d": None}
        seq = input_ids.clone()
        scores_out, logits_out = [], []
        t["prefill_s"] = 0.0
        t["decode_s"] = 0.0
        for step in range(max_new_tokens):
            ts = time.perf_counter()
            first = step == 0
            last_tok = seq[:, -1:]
            # main stream (vcd_sample.py:106-119)
            if first:
                if main_embeds is not None:
                    l, caches["m"] = self._forward(None, inputs_embeds=main_embeds)
                else:
                    l, caches["m"] = self._forward(None, input_ids=seq)
            else:
                l, caches["m"] = self._forward(caches["m"], input_ids=last_tok)
            u = d = None
            if vdd and cd_embeds is not None:
                # model_kwargs_cd = model_kwargs.copy() EVERY step (vcd_sample.py:149): from step 1 on it carries the main
                # stream's cache, so the contrast forward is the main forward again (Appendix A Q7)
                if first:
                    u, _ = self._forward(None, inputs_embeds=cd_embeds)
                else:
                    u = l.clone()
                if use_dd and use_dd_unk:  # vcd_sample.py:171-185: its own persistent kwargs => a real stream with its own cache
                    d, caches["d"] = self._forward(caches["d"], input_ids=debias_ids(seq, "none") if first else last_tok)
            elif vdd:
                # vcd_sample.py:151-169: the first contrast stream is 'unk' when use_dd_unk is set, else 'none'
                if use_dd_unk:
                    u, caches["u"] = self._forward(caches["u"], input_ids=debias_ids(seq, "unk") if first else last_tok)
                if use_dd:
                    # vcd_sample.py:171-185 (third forward when both flags) / :156-160 (only use_dd)
                    d, caches["d"] = self._forward(caches["d"], input_ids=debias_ids(seq, "none") if first else last_tok)
            q = None if q_list is None else q_list[step]
            if do_sample:
                tok, s = vdd_step(l, u, d, cd_alpha, cd_beta, temperature, top_k, top_p, q)
            else:
                s = l
                tok = torch.argmax(l, dim=-1)
            if output_scores:
                scores_out.append(s)
            if keep_stream_logits:
                logits_out.append({"m": l, "u": u, "d": d})
            if eos is not None:
                tok = tok * unfinished + pad * (1 - unfinished)  # vcd_sample.py:257-260
            seq = torch.cat([seq, tok[:, None]], dim=-1)          # vcd_sample.py:263
            if eos is not None:
                unfinished = unfinished * (tok != eos).long()      # vcd_sample.py:285-288
            t["prefill_s" if first else "decode_s"] += time.perf_counter() - ts
            if eos is not None and int(unfinished.max()) == 0:     # vcd_sample.py:291
                break
        t["steps"] = seq.shape[1] - input_ids.shape[1]
        return OracleOutput(seq, scores_out if output_scores else None,
                            logits_out if keep_stream_logits else None, t)
