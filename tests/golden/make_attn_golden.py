"""Golden `attentions[step][layer]` of the UNMODIFIED reference (vcd_sample.py:241-246: the main stream's decoder attentions,
one tuple of per-layer tensors per step) on the tiny config, for the full attention capture of the engine.

Run in a container that has /root/reference:   python tests/golden/make_attn_golden.py
Writes tests/golden/ref_tiny_attn.npz (+ .json).  Same model, inputs and seeds as make_golden.py's cases of the same name, so the
token sequences must reproduce tests/golden/ref_tiny.npz (asserted below).
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, HERE)

import make_golden as G  # noqa: E402

CASES = ("both_greedy", "unk_batch2_greedy")


def main():
    from transformers.generation.logits_process import LogitsProcessorList, TopKLogitsWarper
    from llava_align_b200.config import tiny_config
    from llava_align_b200.weights import make_state_dict
    ref_sample, ref_llama = G.load_reference()
    cfg = tiny_config()
    sd = make_state_dict(cfg, seed=0)
    pinned = np.load(os.path.join(HERE, "ref_tiny.npz"))
    out, meta = {}, []
    for i, (name, B, T, n_new, flags, sampler, eos) in enumerate(G.CASES):
        if name not in CASES:
            continue
        ids, img = G.make_inputs(cfg, B, T, seed=100 + i)
        model = G.build_reference_model(ref_llama, cfg, sd)  # fresh: recording attentions perturbs later forwards (make_golden.py)
        warpers = LogitsProcessorList([TopKLogitsWarper(top_k=sampler["top_k"], min_tokens_to_keep=1)])

        class MaxLength431(list):
            def __call__(self, input_ids, scores, **kw):
                return input_ids.shape[-1] >= T + n_new
        torch.manual_seed(7)
        with torch.inference_mode():
            o = ref_sample.sample(model, ids.clone(), logits_warper=warpers, stopping_criteria=MaxLength431(),
                                  pad_token_id=model.config.pad_token_id, eos_token_id=eos, output_scores=True, output_attentions=True,
                                  return_dict_in_generate=True, attention_mask=torch.ones_like(ids), images=img, use_cache=True, **dict(flags))
        assert np.array_equal(o.sequences.numpy(), pinned[f"{name}.sequences"]), name
        att = o["attentions"]
        assert len(att) == n_new and len(att[0]) == cfg.num_hidden_layers
        for s, per_layer in enumerate(att):
            out[f"{name}.s{s}"] = torch.stack([a.float() for a in per_layer], dim=0).numpy()  # [L, B, heads, q_len, kv_len]
        meta.append(dict(name=name, steps=n_new, layers=cfg.num_hidden_layers,
                         shapes=[list(out[f"{name}.s{s}"].shape) for s in range(n_new)]))
        print(name, [tuple(out[f"{name}.s{s}"].shape) for s in (0, 1, n_new - 1)])
    np.savez_compressed(os.path.join(HERE, "ref_tiny_attn.npz"), **out)
    with open(os.path.join(HERE, "ref_tiny_attn.json"), "w") as f:
        json.dump(dict(cases=meta, generator="tests/golden/make_attn_golden.py", torch=torch.__version__), f, indent=1)


if __name__ == "__main__":
    main()
