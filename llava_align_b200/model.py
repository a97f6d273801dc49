"""The reference's model-level surface, backed by libvdd.

`LlavaLlamaForCausalLM.generate(input_ids, images=..., use_dd=..., use_dd_unk=..., cd_alpha=..., cd_beta=..., ...)`
keeps the argument names, defaults and return layouts the reference's drivers rely on
(reference experiments/eval/sampling/llava_sampling.py:96-109, eval/calibrate/llava_calibrate.py:160-185):

  * VDD is active only after `evolve_vcd_sampling()` (vcd_utils/vcd_sample.py:325-326) and only when
    `do_sample=True` (do_sample=False dispatches to HF greedy_search, which the reference does not patch);
  * it is switched on by `use_dd`, `use_dd_unk` (vcd_sample.py:122-124,147); `cd_alpha` defaults to 0.5 and
    `cd_beta` to 0.1 when None (vcd_sample.py:188-189);
  * warpers follow HF-4.31 `_get_logits_warper`: a kwarg that is passed (even as None) overrides
    `generation_config`, Temperature -> TopK -> TopP;
  * the result starts with the untouched input_ids, -200 sentinel included (llava_sampling.py:111-114).
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Optional

import torch

from . import engine as _engine
from .config import VddConfig, IMAGE_TOKEN_INDEX

_VDD_SAMPLER_INSTALLED = False


def evolve_vcd_sampling():
    """Install the VDD sampler (reference vcd_utils/vcd_sample.py:325-326).  Zero-arg, idempotent."""
    global _VDD_SAMPLER_INSTALLED
    _VDD_SAMPLER_INSTALLED = True


def vdd_sampler_installed() -> bool:
    return _VDD_SAMPLER_INSTALLED


class GenerateOutput(dict):
    """Mapping with attribute access, like HF's SampleDecoderOnlyOutput (vcd_sample.py:316-321)."""
    __getattr__ = dict.get


class _StepAttentions(tuple):
    """`attentions[step]` when only the map the reference's drivers read was materialised -- step 0, LAST layer
    (`attentions[0][-1]`, llava_calibrate.py:180, MME/run_llava.py:218), computed after the prefill from the last layer's
    queries and K cache.  Any other entry raises instead of returning something made up.  The full tuple (every step, every
    layer) is returned instead when it fits `VDD_ATTENTIONS_MAX_GB` (default 4) or `VDD_ATTENTIONS=all` is set."""

    def __new__(cls, step, num_layers, last_layer_map):
        self = super().__new__(cls, ())
        self._step, self._L, self._map = step, num_layers, last_layer_map
        return self

    def __len__(self):
        return self._L

    def __getitem__(self, i):
        if isinstance(i, int) and (i == -1 or i == self._L - 1) and self._map is not None:
            return self._map
        raise NotImplementedError(f"attentions[{self._step}][{i}] was not materialised: this generate kept only attentions[0][-1] "
                                  f"(step 0, last layer, main stream) because the full tuple exceeds VDD_ATTENTIONS_MAX_GB; "
                                  f"set VDD_ATTENTIONS=all to capture every step and layer")

    def __iter__(self):
        raise NotImplementedError("only attentions[0][-1] was materialised (set VDD_ATTENTIONS=all for the full tuple)")


def _attention_capture_mode(B, heads, L, len0, max_new):
    """'all' when the whole `attentions` tuple (vcd_sample.py:241-246) fits the budget, else 'last' (what the drivers read)."""
    import os
    mode = os.environ.get("VDD_ATTENTIONS", "auto")
    if mode in ("all", "last"):
        return mode
    total = 4 * B * heads * L * (len0 * len0 + max(max_new - 1, 0) * (len0 + max_new - 1))
    return "all" if total <= float(os.environ.get("VDD_ATTENTIONS_MAX_GB", "4")) * 2 ** 30 else "last"


class _VisionTowerHandle:
    """What drivers touch on `model.get_vision_tower()` (llava_sampling.py:49-53)."""

    def __init__(self, image_processor=None):
        self.is_loaded = True
        self.image_processor = image_processor

    def load_model(self):
        self.is_loaded = True

    def to(self, *a, **k):
        return self


class LlavaLlamaForCausalLM:
    def __init__(self, cfg: VddConfig, device="cuda:0", tp_size=1, tp_rank=0, image_processor=None):
        self.vdd_config = cfg
        self.engine = _engine.VddEngine(cfg, device, tp_size, tp_rank)
        self.device = self.engine.device
        self.dtype = torch.bfloat16
        self.config = SimpleNamespace(mm_use_im_start_end=False, mm_use_im_patch_token=False, vocab_size=cfg.vocab_size,
                                      hidden_size=cfg.hidden_size, pad_token_id=cfg.pad_token_id,
                                      eos_token_id=cfg.eos_token_id, bos_token_id=cfg.bos_token_id,
                                      max_position_embeddings=cfg.max_position_embeddings, is_encoder_decoder=False)
        # no `max_sequence_length`: LLaVA-1.5 configs do not define it, so the reference's loader reports context_len 2048
        # (builder.py:143-146); the checkpoint loader sets it when a config.json does
        # HF GenerationConfig defaults (temperature 1.0, top_k 50, top_p 1.0)
        self.generation_config = SimpleNamespace(temperature=1.0, top_k=50, top_p=1.0, do_sample=False, max_new_tokens=None,
                                                 max_length=20, eos_token_id=cfg.eos_token_id, pad_token_id=cfg.pad_token_id,
                                                 output_scores=False, output_attentions=False, return_dict_in_generate=False)
        self._vt = _VisionTowerHandle(image_processor)

    # ---- driver-facing odds and ends
    def get_vision_tower(self):
        return self._vt

    def resize_token_embeddings(self, n):
        if n != self.vdd_config.vocab_size:
            raise NotImplementedError("resizing the vocabulary of a loaded engine is not supported")

    def eval(self):
        return self

    def cuda(self, *a, **k):
        return self

    def to(self, *a, **k):
        return self

    def load_state_dict(self, sd, strict=True):
        self.engine.load_state_dict(sd)

    # ---- generate
    @torch.no_grad()
    def generate(self, input_ids=None, images=None, images_cd=None, cd_alpha=None, cd_beta=None, use_dd=None,
                 use_dd_unk=None, max_new_tokens=None, max_length=None, use_cache=True, num_beams=1, stopping_criteria=None,
                 attention_mask=None, **kwargs):
        if input_ids is None:
            input_ids = kwargs.pop("inputs")
        gc = self.generation_config
        opt = {k: kwargs.pop(k) if k in kwargs else getattr(gc, k) for k in
               ("do_sample", "temperature", "top_k", "top_p", "output_scores", "output_attentions",
                "return_dict_in_generate", "eos_token_id", "pad_token_id")}
        kwargs.pop("output_hidden_states", None)
        # extension (not a reference kwarg): also return the debias streams' own first-step scores -- what the post-hoc-debias
        # drivers get from two extra generates per question (llava_calibrate.py:192-194), folded into this fused batch
        want_stream_scores = bool(kwargs.pop("return_stream_scores", False))
        if kwargs:
            raise ValueError(f"The following `model_kwargs` are not used by the model: {sorted(kwargs)}")  # HF _validate_model_kwargs
        if num_beams != 1:
            raise NotImplementedError("beam search is outside the VDD path (reference scripts use num_beams=1)")
        if stopping_criteria is not None and len(stopping_criteria) > 0:
            # the reference's drivers build a KeywordsStoppingCriteria but never pass it (llava_sampling.py:91-109)
            raise NotImplementedError("custom stopping_criteria are not supported: EOS and max_new_tokens stop generation")
        if input_ids.dim() == 1:
            input_ids = input_ids.unsqueeze(0)
        ragged = None
        if attention_mask is not None and not bool(torch.as_tensor(attention_mask).bool().all()):
            # Ragged batch (designed, not copied: the reference rebuilds the decode-step mask as all ones, llava_arch.py:91-94,
            # SURVEY Appendix A Q8, so its padded batches attend to pad tokens).  Masked positions -- left or right padding --
            # do not exist for any stream: every row computes what the same prompt computes alone.
            ragged = torch.as_tensor(attention_mask)
            if tuple(ragged.shape) != tuple(input_ids.shape):
                raise ValueError(f"attention_mask {tuple(ragged.shape)} does not match input_ids {tuple(input_ids.shape)}")
            if bool(((input_ids == IMAGE_TOKEN_INDEX) & (ragged.to(input_ids.device) == 0)).any()):
                raise ValueError("the image sentinel of a row is masked out")
        B, T = input_ids.shape
        if max_new_tokens is None:
            max_new_tokens = gc.max_new_tokens
        if max_new_tokens is None:
            max_new_tokens = (max_length or gc.max_length) - T
        if max_new_tokens < 1:
            raise ValueError("max_new_tokens must be >= 1")
        eos = opt["eos_token_id"]
        if isinstance(eos, (list, tuple)):
            if len(eos) > 1:
                raise NotImplementedError("a single eos_token_id is supported")
            eos = eos[0] if eos else None
        pad = opt["pad_token_id"]
        if eos is not None and pad is None:
            raise ValueError("If `eos_token_id` is defined, make sure that `pad_token_id` is defined.")  # vcd_sample.py:258-259

        has_image = images is not None and bool((input_ids == IMAGE_TOKEN_INDEX).any())
        do_sample = bool(opt["do_sample"])
        # VDD/VCD is active iff images_cd is given or use_dd / use_dd_unk is set (vcd_sample.py:122-124,147)
        use_cd = images_cd is not None and has_image
        vdd = do_sample and vdd_sampler_installed() and bool(use_dd or use_dd_unk or use_cd)
        # images_cd together with the debias flags: the engine applies the reference's precedence (vcd_sample.py:148 -- the cd
        # forward replaces the unk stream; :171-185 -- the 'none' stream joins only when both flags are set)
        use_unk = bool(vdd and use_dd_unk)
        use_none = bool(vdd and use_dd)
        if do_sample:
            t, k, p = opt["temperature"], opt["top_k"], opt["top_p"]
            if t is not None and not t > 0:
                raise ValueError("`temperature` has to be a strictly positive float")
            params = _engine.make_sample_params(
                cd_alpha=0.5 if cd_alpha is None else cd_alpha, cd_beta=0.1 if cd_beta is None else cd_beta,
                temperature=None if (t is None or t == 1.0) else t, top_k=None if (k is None or k == 0) else k,
                top_p=None if (p is None or p >= 1.0) else p, use_contrast=vdd, eos_token_id=eos, pad_token_id=pad)
        else:  # greedy_search: argmax of the main stream, no VDD
            params = _engine.make_sample_params(use_contrast=False, eos_token_id=eos, pad_token_id=pad)
        V = self.vdd_config.vocab_size
        q = None
        gen = None
        if do_sample:
            # torch.multinomial(probs, 1) == argmax(probs / q), q = empty_like(probs).exponential_(1) from torch's CUDA generator.
            # K8 regenerates exactly those numbers from the generator's (seed, offset) -- nothing of size [steps, B, V] is
            # allocated -- and the generator is advanced afterwards as if the reference's loop had drawn them.
            # VDD_TORCH_Q=1: draw them with torch up front instead (one [steps, B, V] call: a different stream of numbers).
            import os
            if os.environ.get("VDD_TORCH_Q"):
                if params.top_k != 1:
                    q = torch.empty(max_new_tokens, B, V, dtype=torch.float32, device=self.device).exponential_(1)
                    q = self.engine.tp_broadcast(q)  # tensor parallel: every rank must race against rank 0's variates
            else:
                gen = torch.cuda.default_generators[self.device.index if self.device.index is not None else torch.cuda.current_device()]
                s64 = gen.initial_seed() & (2 ** 64 - 1)
                state = torch.tensor([s64 - 2 ** 64 if s64 >= 2 ** 63 else s64, gen.get_offset()], dtype=torch.int64, device=self.device)
                state = self.engine.tp_broadcast(state)  # tensor parallel: rank 0's generator decides for all ranks
                seed, offset = (int(x) for x in state.tolist())
                seed &= 2 ** 64 - 1
                self.engine.set_sampler_rng(seed, offset)
        want_scores = bool(opt["output_scores"] and opt["return_dict_in_generate"])
        scores_buf = torch.empty(max_new_tokens, B, V, dtype=torch.float32, device=self.device) if want_scores else None
        attn_buf = attn_all = None
        if opt["output_attentions"] and opt["return_dict_in_generate"]:
            if ragged is not None:
                raise NotImplementedError("attention maps need rows of equal length")
            Tm = T + (self.vdd_config.num_patches - 1 if has_image else 0)
            nh, nl = self.vdd_config.num_attention_heads, self.vdd_config.num_hidden_layers
            if _attention_capture_mode(B, nh, nl, Tm, max_new_tokens) == "all":
                attn_all = (torch.empty(nl, B, nh, Tm, Tm, dtype=torch.float32, device=self.device),
                            torch.zeros(max_new_tokens - 1, nl, B, nh, Tm + max_new_tokens - 1, dtype=torch.float32, device=self.device)
                            if max_new_tokens > 1 else None)
            else:
                attn_buf = torch.empty(B, nh, Tm, Tm, dtype=torch.float32, device=self.device)
        stream_names = []
        if vdd and use_cd:
            stream_names = (["none"] if (use_unk and use_none) else []) + ["cd"]  # engine precedence: cd replaces unk (vcd_sample.py:148)
        elif vdd:
            stream_names = (["unk"] if use_unk else []) + (["none"] if use_none else [])
        ss_buf = None
        if want_stream_scores and stream_names and opt["return_dict_in_generate"]:
            ss_buf = torch.empty(len(stream_names), B, V, dtype=torch.float32, device=self.device)
        new_tokens, steps = self.engine.generate(input_ids, images if has_image else None, use_unk, use_none, params,
                                                 max_new_tokens, q, scores_buf, images_cd if (vdd and use_cd) else None, attn_buf, ss_buf,
                                                 ragged, attn_all)
        if gen is not None:
            gen.set_offset(offset + self.engine.sampler_rng_advance())
        seq = torch.cat([input_ids.to("cpu", torch.int64), new_tokens], dim=1).to(input_ids.device)
        if not opt["return_dict_in_generate"]:
            return seq
        out = GenerateOutput(sequences=seq)
        out["scores"] = tuple(scores_buf[i] for i in range(steps)) if want_scores else None
        # vcd_sample.py:241-246: decoder_attentions of the MAIN stream, one entry per step
        L = self.vdd_config.num_hidden_layers
        if attn_all is not None:
            # HF's shape: step 0 -> L tensors [B, heads, len, len]; step s -> L tensors [B, heads, 1, len + s]
            a0, an = attn_all
            Tm = a0.shape[-1]
            out["attentions"] = tuple(tuple(a0[l] for l in range(L)) if i == 0 else
                                      tuple(an[i - 1, l, :, :, None, :Tm + i] for l in range(L)) for i in range(steps))
        else:
            out["attentions"] = (tuple(_StepAttentions(i, L, attn_buf if i == 0 else None) for i in range(steps))
                                 if attn_buf is not None else None)
        out["hidden_states"] = None
        if ss_buf is not None:
            out["stream_scores"] = {n: ss_buf[i] for i, n in enumerate(stream_names)}
        return out
