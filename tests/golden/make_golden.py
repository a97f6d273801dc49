"""Generate tests/golden/ref_*.npz by running the UNMODIFIED reference code on tiny seeded models.

Run in the authoring container only (needs /root/reference):  python tests/golden/make_golden.py

What executes from /root/reference, untouched:
  vcd_utils/vcd_sample.py::sample                      (the patched sampling loop)
  experiments/llava/model/language_model/llava_llama.py (LlavaLlamaForCausalLM.forward, prepare_inputs_for_generation[_cd])
  experiments/llava/model/llava_arch.py                 (prepare_inputs_labels_for_multimodal, encode_images)
  experiments/llava/model/multimodal_encoder/clip_encoder.py, multimodal_projector/builder.py

Shims (the reference targets transformers 4.31; 5.5.0 is installed here) -- names only, no arithmetic:
  * transformers.generation.utils.SampleOutput / transformers.generation.Sample*Output: output dataclasses
  * package stubs for `experiments`, `experiments.llava`, `llava`, `llava.model`, `vcd_utils` so their
    __init__ files (which import MPT/LAVIS) are not executed
  * AutoConfig.register('llava') is a no-op (transformers 5.5 already owns a 'llava' model type)
  * DynamicCache.__getitem__ -> (keys, values) of a layer, for the tuple-cache indexing at llava_arch.py:93
  * a 4.31-style max-length stopping criterion returning a python bool (vcd_sample.py:295)
  * the model's forward deep-copies `past_key_values` on entry: transformers-4.31 caches are immutable tuples (every
    forward returns NEW tensors via torch.cat), 5.5's DynamicCache is updated in place.  The reference relies on the
    former when it hands the main stream's cache to the use_cd contrast forward (`model_kwargs.copy()` every step,
    vcd_sample.py:148-150); without the copy that forward would append the token to the main cache a second time.
`sample()` is called directly the way transformers-4.31 `generate()` calls it: logits_warper built in
_get_logits_warper order (Temperature, TopK, TopP), MaxLengthCriteria, attention_mask of ones, and the
drivers' kwargs (images, use_dd, use_dd_unk, cd_alpha, cd_beta, use_cache) as model_kwargs.
"""
import os
import sys
import tempfile
import types
from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference"
sys.path.insert(0, REPO)


def load_reference():
    import transformers  # noqa
    import transformers.generation as gen
    import transformers.generation.utils as gu
    from transformers.utils import ModelOutput

    @dataclass
    class SampleDecoderOnlyOutput(ModelOutput):
        sequences: torch.LongTensor = None
        scores: Optional[Tuple[torch.FloatTensor]] = None
        attentions: Optional[Tuple] = None
        hidden_states: Optional[Tuple] = None

    @dataclass
    class SampleEncoderDecoderOutput(ModelOutput):
        sequences: torch.LongTensor = None

    gu.SampleOutput = SampleDecoderOnlyOutput
    gen.SampleDecoderOnlyOutput = SampleDecoderOnlyOutput
    gen.SampleEncoderDecoderOutput = SampleEncoderDecoderOutput

    def pkg(name, path):
        m = types.ModuleType(name)
        m.__path__ = [path]
        sys.modules[name] = m

    pkg("experiments", REF + "/experiments")
    pkg("experiments.llava", REF + "/experiments/llava")
    pkg("llava", REF + "/experiments/llava")
    pkg("llava.model", REF + "/experiments/llava/model")
    pkg("vcd_utils", REF + "/vcd_utils")
    from transformers import AutoConfig, AutoModelForCausalLM
    AutoConfig.register = staticmethod(lambda *a, **k: None)
    AutoModelForCausalLM.register = classmethod(lambda cls, *a, **k: None)
    from transformers.cache_utils import DynamicCache

    def _getitem(self, i):
        layer = self.layers[i]
        return (layer.keys, layer.values)

    DynamicCache.__getitem__ = _getitem
    import vcd_utils.vcd_sample as ref_sample
    import llava.model.language_model.llava_llama as ref_llama
    return ref_sample, ref_llama


def build_reference_model(ref_llama, cfg, sd):
    from transformers import CLIPVisionConfig, CLIPVisionModel, CLIPImageProcessor
    tmp = tempfile.mkdtemp()
    vc = CLIPVisionConfig(hidden_size=cfg.vision_hidden_size, intermediate_size=cfg.vision_intermediate_size,
                          num_hidden_layers=cfg.vision_num_hidden_layers,
                          num_attention_heads=cfg.vision_num_attention_heads, image_size=cfg.image_size,
                          patch_size=cfg.patch_size, layer_norm_eps=cfg.vision_layer_norm_eps)
    CLIPVisionModel(vc).save_pretrained(tmp)
    CLIPImageProcessor(size={"shortest_edge": cfg.image_size},
                       crop_size={"height": cfg.image_size, "width": cfg.image_size}).save_pretrained(tmp)
    lcfg = ref_llama.LlavaConfig(
        hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size,
        num_hidden_layers=cfg.num_hidden_layers, num_attention_heads=cfg.num_attention_heads,
        num_key_value_heads=cfg.num_attention_heads, vocab_size=cfg.vocab_size,
        max_position_embeddings=cfg.max_position_embeddings, rms_norm_eps=cfg.rms_norm_eps,
        rope_theta=cfg.rope_theta, attn_implementation="eager", tie_word_embeddings=False,
        mm_vision_tower=tmp, mm_hidden_size=cfg.vision_hidden_size, mm_projector_type="mlp2x_gelu",
        mm_vision_select_layer=cfg.mm_vision_select_layer, mm_vision_select_feature="patch",
        pad_token_id=cfg.pad_token_id, bos_token_id=cfg.bos_token_id, eos_token_id=cfg.eos_token_id)
    model = ref_llama.LlavaLlamaForCausalLM(lcfg).eval()
    model.get_vision_tower().load_model()
    res = model.load_state_dict(sd, strict=False)
    bad = [k for k in res.missing_keys if "rotary" not in k and "position_ids" not in k]
    assert not bad and not res.unexpected_keys, (bad, res.unexpected_keys)
    import copy
    orig_forward = model.forward

    def forward_431(*a, **k):  # immutable-cache semantics of transformers 4.31 (see the module docstring)
        if k.get("past_key_values") is not None:
            k["past_key_values"] = copy.deepcopy(k["past_key_values"])
        return orig_forward(*a, **k)

    model.forward = forward_431
    return model


def make_inputs(cfg, B, T, seed):
    """Same synthetic inputs the tests rebuild: ids uniform in [3,V), one -200 per row at 1+T//8; image N(0,1)
    rounded to bf16."""
    rng = np.random.default_rng(seed)
    ids = rng.integers(3, cfg.vocab_size, size=(B, T)).astype(np.int64)
    ids[:, 1 + T // 8] = -200
    img = rng.standard_normal((B, 3, cfg.image_size, cfg.image_size), dtype=np.float32)
    img = torch.from_numpy(img).to(torch.bfloat16).float()
    return torch.from_numpy(ids), img


ATTN_CASES = ("both_greedy", "unk_batch2_greedy", "cd_both_greedy")  # cases that also pin attentions[0][-1]


def run_reference(ref_sample, model, ids, img, n_new, flags, sampler, eos, seed, img_cd=None, want_attn=False):
    from transformers.generation.logits_process import (LogitsProcessorList, TemperatureLogitsWarper,
                                                        TopKLogitsWarper, TopPLogitsWarper)

    class MaxLength431(list):
        """transformers-4.31 StoppingCriteriaList([MaxLengthCriteria(n)]) returned a python bool
        (`input_ids.shape[-1] >= max_length`); 5.5 returns a per-row tensor, which vcd_sample.py:295 cannot
        branch on for B>1."""
        def __init__(self, max_length):
            super().__init__()
            self.max_length = max_length

        def __call__(self, input_ids, scores, **kw):
            return input_ids.shape[-1] >= self.max_length
    warpers = LogitsProcessorList()
    t, k, p = sampler.get("temperature"), sampler.get("top_k"), sampler.get("top_p")
    if t is not None and t != 1.0:
        warpers.append(TemperatureLogitsWarper(t))
    if k is not None and k != 0:
        warpers.append(TopKLogitsWarper(top_k=k, min_tokens_to_keep=1))
    if p is not None and p < 1.0:
        warpers.append(TopPLogitsWarper(top_p=p, min_tokens_to_keep=1))
    torch.manual_seed(seed)
    with torch.inference_mode():
        out = ref_sample.sample(
            model, ids.clone(), logits_warper=warpers,
            stopping_criteria=MaxLength431(ids.shape[1] + n_new),
            pad_token_id=model.config.pad_token_id, eos_token_id=eos, output_scores=True, output_attentions=want_attn,
            return_dict_in_generate=True, attention_mask=torch.ones_like(ids), images=img, images_cd=img_cd,
            use_cache=True, **flags)
    if want_attn:
        # what llava_calibrate.py:180 reads: step 0, last layer, main stream -> [B, heads, len, len]
        return out.sequences, torch.stack(out.scores, dim=0), out["attentions"][0][-1].float()
    return out.sequences, torch.stack(out.scores, dim=0)  # [B,T+n], [n,B,V]


CASES = [
    # name, B, T, n_new, flags, sampler, eos
    ("dd_greedy", 1, 12, 8, dict(use_dd=True, use_dd_unk=False, cd_alpha=1.0, cd_beta=0.1), dict(top_k=1), None),
    ("unk_greedy", 1, 12, 8, dict(use_dd=False, use_dd_unk=True, cd_alpha=1.0, cd_beta=0.1), dict(top_k=1), None),
    ("both_greedy", 1, 12, 8, dict(use_dd=True, use_dd_unk=True, cd_alpha=1.0, cd_beta=0.1), dict(top_k=1), None),
    ("both_defaults_sample", 1, 16, 8, dict(use_dd=True, use_dd_unk=True), dict(), None),
    ("both_t07_k5", 1, 12, 8, dict(use_dd=True, use_dd_unk=True, cd_alpha=1.0, cd_beta=0.2),
     dict(temperature=0.7, top_k=5), None),
    ("unk_p09", 1, 12, 8, dict(use_dd=False, use_dd_unk=True, cd_alpha=0.5, cd_beta=0.01), dict(top_p=0.9), None),
    ("plain_sample_t1", 1, 12, 8, dict(use_dd=False, use_dd_unk=False), dict(top_p=0.8, top_k=50), None),
    ("unk_batch2_greedy", 2, 12, 8, dict(use_dd=False, use_dd_unk=True, cd_alpha=1.0, cd_beta=0.1), dict(top_k=1), None),
    ("unk_batch2_eos", 2, 12, 8, dict(use_dd=False, use_dd_unk=True, cd_alpha=1.0, cd_beta=0.1), dict(top_k=1), "auto"),
    # use_cd: images_cd = add_diffusion_noise(image, 500) (vcd_utils/vcd_add_noise.py), VCD baseline of the reference
    ("cd_greedy", 1, 12, 8, dict(use_dd=False, use_dd_unk=False, cd_alpha=1.0, cd_beta=0.1, use_cd=500), dict(top_k=1), None),
    ("cd_sample_batch2", 2, 12, 8, dict(use_dd=False, use_dd_unk=False, cd_alpha=0.5, cd_beta=0.2, use_cd=700), dict(top_p=0.9), None),
    # images_cd together with the debias flags (vcd_sample.py:148 wins the first contrast forward; :171-185 adds 'none' when both are set)
    ("cd_both_greedy", 1, 12, 8, dict(use_dd=True, use_dd_unk=True, cd_alpha=1.0, cd_beta=0.1, use_cd=500), dict(top_k=1), None),
    ("cd_both_sample", 1, 12, 8, dict(use_dd=True, use_dd_unk=True, cd_alpha=0.5, cd_beta=0.1, use_cd=600), dict(temperature=0.8, top_k=20), None),
    ("cd_unk_only_greedy", 1, 12, 8, dict(use_dd=False, use_dd_unk=True, cd_alpha=1.0, cd_beta=0.1, use_cd=500), dict(top_k=1), None),
]


def main():
    from llava_align_b200.config import tiny_config
    from llava_align_b200.weights import make_state_dict
    ref_sample, ref_llama = load_reference()
    cfg = tiny_config()
    sd = make_state_dict(cfg, seed=0)
    model = build_reference_model(ref_llama, cfg, sd)
    out = {}
    meta = []
    for i, (name, B, T, n_new, flags, sampler, eos) in enumerate(CASES):
        ids, img = make_inputs(cfg, B, T, seed=100 + i)
        flags = dict(flags)
        img_cd = None
        if flags.get("use_cd"):
            sys.path.insert(0, REF)
            from vcd_utils.vcd_add_noise import add_diffusion_noise  # the reference's own noising, unmodified
            torch.manual_seed(1000 + i)
            img_cd = add_diffusion_noise(img, flags.pop("use_cd")).to(torch.bfloat16).float()
            out[f"{name}.images_cd"] = img_cd.numpy()
        if eos == "auto":
            # pick the token row 0 emits at step 2 as EOS so rows finish at different steps
            seq0, _ = run_reference(ref_sample, model, ids, img, n_new, flags, sampler, None, seed=7)
            eos = int(seq0[0, T + 2])
        seq, scores = run_reference(ref_sample, model, ids, img, n_new, flags, sampler, eos, seed=7, img_cd=img_cd)
        if name in ATTN_CASES:
            # Second run with output_attentions=True on a FRESH model: under transformers 5.5 a forward that records attentions
            # leaves state behind that perturbs later forwards of the same module (observed: the next run's tokens change),
            # so it must not share a model with the other cases.  Its own tokens must reproduce the first run.
            fresh = build_reference_model(ref_llama, cfg, sd)
            seq_a, _, attn = run_reference(ref_sample, fresh, ids, img, n_new, flags, sampler, eos, seed=7, img_cd=img_cd, want_attn=True)
            assert torch.equal(seq_a, seq), name
            out[f"{name}.attn0_last"] = attn.numpy()
        out[f"{name}.sequences"] = seq.numpy()
        out[f"{name}.scores"] = scores.float().numpy()
        meta.append(dict(name=name, B=B, T=T, n_new=n_new, flags=dict(flags, use_cd=img_cd is not None), sampler=sampler, eos=eos,
                         input_seed=100 + i, sample_seed=7))
        print(name, seq[:, T:].tolist())
    import json
    np.savez_compressed(os.path.join(REPO, "tests/golden/ref_tiny.npz"), **out)
    with open(os.path.join(REPO, "tests/golden/ref_tiny.json"), "w") as f:
        json.dump(dict(config=cfg.to_dict(), weight_seed=0, cases=meta,
                       generator="tests/golden/make_golden.py", torch=torch.__version__), f, indent=1)


if __name__ == "__main__":
    main()
