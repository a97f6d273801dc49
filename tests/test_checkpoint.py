"""Loader for an on-disk HF LLaVA-1.5 checkpoint (reference experiments/llava/model/builder.py:97-103,137-141): no real
checkpoint can be fetched offline, so a synthetic one is written in the released layout -- LLaVA shards without the CLIP
tower, `mm_vision_tower` pointing at a CLIP directory with `vision_model.*` keys, plus the non-parameter keys real
checkpoints carry -- and read back."""
import json
import os

import pytest
import torch

from helpers import make_inputs, tiny_model


def write_checkpoint(root, cfg, sd, fmt="safetensors"):
    llava, clip = os.path.join(root, "llava-tiny"), os.path.join(root, "clip-tiny")
    os.makedirs(llava)
    os.makedirs(clip)
    vt = "model.vision_tower.vision_tower.vision_model."
    lm = {k: v.to(torch.bfloat16) for k, v in sd.items() if not k.startswith(vt)}  # exact: the synthetic weights are bf16 values
    vis = {"vision_model." + k[len(vt):]: v.to(torch.bfloat16) for k, v in sd.items() if k.startswith(vt)}
    # keys real checkpoints carry that are not parameters of the path
    lm["model.layers.0.self_attn.rotary_emb.inv_freq"] = torch.ones(64)
    vis["vision_model.embeddings.position_ids"] = torch.arange(cfg.num_patches + 1).unsqueeze(0)
    vis["text_model.embeddings.token_embedding.weight"] = torch.zeros(4, 4, dtype=torch.float16)
    vis["logit_scale"] = torch.tensor(2.6)
    json.dump(dict(architectures=["LlavaLlamaForCausalLM"], model_type="llava", hidden_size=cfg.hidden_size,
                   intermediate_size=cfg.intermediate_size, num_hidden_layers=cfg.num_hidden_layers,
                   num_attention_heads=cfg.num_attention_heads, num_key_value_heads=cfg.num_attention_heads,
                   vocab_size=cfg.vocab_size, max_position_embeddings=cfg.max_position_embeddings, rms_norm_eps=cfg.rms_norm_eps,
                   rope_theta=cfg.rope_theta, mm_vision_tower=clip, mm_hidden_size=cfg.vision_hidden_size,
                   mm_projector_type="mlp2x_gelu", mm_vision_select_layer=-2, mm_vision_select_feature="patch",
                   mm_use_im_start_end=False, pad_token_id=0, bos_token_id=1, eos_token_id=2, _name_or_path="tiny"),
              open(os.path.join(llava, "config.json"), "w"))
    json.dump(dict(model_type="clip", vision_config=dict(hidden_size=cfg.vision_hidden_size, intermediate_size=cfg.vision_intermediate_size,
                                                         num_hidden_layers=cfg.vision_num_hidden_layers,
                                                         num_attention_heads=cfg.vision_num_attention_heads, image_size=cfg.image_size,
                                                         patch_size=cfg.patch_size, layer_norm_eps=cfg.vision_layer_norm_eps)),
              open(os.path.join(clip, "config.json"), "w"))
    keys = sorted(lm)
    half = len(keys) // 2
    if fmt == "safetensors":
        from safetensors.torch import save_file
        save_file({k: lm[k].contiguous() for k in keys[:half]}, os.path.join(llava, "model-00001-of-00002.safetensors"))
        save_file({k: lm[k].contiguous() for k in keys[half:]}, os.path.join(llava, "model-00002-of-00002.safetensors"))
        save_file({k: v.contiguous() for k, v in vis.items()}, os.path.join(clip, "model.safetensors"))
    else:
        torch.save({k: lm[k] for k in keys[:half]}, os.path.join(llava, "pytorch_model-00001-of-00002.bin"))
        torch.save({k: lm[k] for k in keys[half:]}, os.path.join(llava, "pytorch_model-00002-of-00002.bin"))
        torch.save(vis, os.path.join(clip, "pytorch_model.bin"))
    return llava, clip


@pytest.mark.parametrize("fmt", ["safetensors", "bin"])
def test_checkpoint_reader_recovers_config_and_every_tensor(tmp_path, fmt):
    from llava_align_b200.checkpoint import config_from_hf, iter_checkpoint_tensors
    from llava_align_b200.weights import state_dict_shapes
    cfg, sd, _ = tiny_model(0)
    llava, _ = write_checkpoint(str(tmp_path), cfg, sd, fmt)
    got_cfg = config_from_hf(llava)
    for f in ("hidden_size", "intermediate_size", "num_hidden_layers", "num_attention_heads", "vocab_size", "rms_norm_eps",
              "vision_hidden_size", "vision_intermediate_size", "vision_num_hidden_layers", "vision_num_attention_heads",
              "image_size", "patch_size", "mm_vision_select_layer"):
        assert getattr(got_cfg, f) == getattr(cfg, f), f
    seen = {}
    for name, t in iter_checkpoint_tensors(llava):
        assert name not in seen
        seen[name] = t
    assert set(seen) == set(state_dict_shapes(cfg)), set(seen) ^ set(state_dict_shapes(cfg))
    for k, v in sd.items():
        assert torch.equal(seen[k].float(), v), k


def test_checkpoint_config_rejects_what_is_not_llava15(tmp_path):
    from llava_align_b200.checkpoint import config_from_hf
    cfg, sd, _ = tiny_model(0)
    llava, _ = write_checkpoint(str(tmp_path), cfg, sd)
    c = json.load(open(os.path.join(llava, "config.json")))
    c["num_key_value_heads"] = 1
    json.dump(c, open(os.path.join(llava, "config.json"), "w"))
    with pytest.raises(NotImplementedError):
        config_from_hf(llava)


@pytest.mark.gpu
def test_load_pretrained_model_from_directory_matches_state_dict(tmp_path):
    """load_pretrained_model(<dir>) (reference builder.py:26,148 signature) == the same weights pushed tensor by tensor."""
    from llava_align_b200.builder import load_pretrained_model
    from llava_align_b200.model import LlavaLlamaForCausalLM, evolve_vcd_sampling
    cfg, sd, _ = tiny_model(0)
    llava, _ = write_checkpoint(str(tmp_path), cfg, sd)
    evolve_vcd_sampling()
    tok, m, proc, ctx = load_pretrained_model(llava, None, "llava-tiny")
    assert tok is None and proc is None and ctx == 2048  # reference builder.py:143-146: no max_sequence_length in the config
    ref = LlavaLlamaForCausalLM(cfg, "cuda:0")
    ref.load_state_dict(sd)
    ids, img = make_inputs(cfg, 2, 12, 31)
    kw = dict(images=img.cuda(), use_dd=False, use_dd_unk=True, cd_alpha=1.0, cd_beta=0.1, do_sample=True, top_k=1, max_new_tokens=6,
              output_scores=True, return_dict_in_generate=True, eos_token_id=None)
    a, b = m.generate(ids.cuda(), **kw), ref.generate(ids.cuda(), **kw)
    assert torch.equal(a["sequences"], b["sequences"])
    assert torch.equal(torch.stack(a["scores"]), torch.stack(b["scores"]))
