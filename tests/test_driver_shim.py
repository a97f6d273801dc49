"""The reference's eval drivers against this repo, unchanged.

CPU part (needs /root/reference, i.e. the authoring container): the UNMODIFIED driver files
experiments/eval/sampling/llava_sampling.py and experiments/eval/calibrate/llava_calibrate.py are imported with this repo's
`llava/` and `vcd_utils/` shims first on sys.path and their `eval_model(args)` is executed end to end; only the loaded model is
a stand-in that binds every `generate(...)` call against the signature of llava_align_b200's LlavaLlamaForCausalLM.generate
(no GPU here), the tokenizer / image processor are small fakes, and `.cuda()` is a no-op.

GPU part: the same call sites (llava_sampling.py:96-114, llava_calibrate.py:160-185 and :62-82) run against the real engine
on the tiny model, with the fallback host-side helpers (no reference tree on the GPU box)."""
import argparse
import importlib.util
import inspect
import json
import os
import sys

import numpy as np
import pytest
import torch

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class FakeTokenizer:
    """Whitespace tokenizer with a BOS, enough for tokenizer_image_token / batch_decode / decode."""
    bos_token_id, eos_token_id, unk_token_id, pad_token_id = 1, 2, 0, 0

    def __init__(self, vocab=512):
        self.vocab = vocab

    def __call__(self, text):
        ids = [self.bos_token_id] + [3 + (sum(map(ord, w)) % (self.vocab - 3)) for w in text.split()]
        return argparse.Namespace(input_ids=ids)

    def decode(self, tok, **kw):
        return "yes" if int(tok) % 2 == 0 else "no"

    def batch_decode(self, ids, skip_special_tokens=True):
        return [" ".join(self.decode(t) for t in row) for row in ids]

    def __len__(self):
        return self.vocab


class FakeProcessor:
    image_mean = [0.48145466, 0.4578275, 0.40821073]

    def __init__(self, size=56):
        self.size = size

    def preprocess(self, image, return_tensors="pt"):
        a = np.asarray(image.convert("RGB").resize((self.size, self.size)), dtype=np.float32) / 255.0
        return {"pixel_values": [torch.from_numpy(a).permute(2, 0, 1).contiguous()]}


def _purge(prefixes=("llava", "vcd_utils")):
    for k in list(sys.modules):
        if any(k == p or k.startswith(p + ".") for p in prefixes) and not k.startswith("llava_align_b200"):
            del sys.modules[k]


def _load_driver(path, name, monkeypatch, with_reference_helpers=True):
    if with_reference_helpers:
        monkeypatch.setenv("LLAVA_ALIGN_REFERENCE_LLAVA", os.path.join(REF, "experiments", "llava"))
    else:
        monkeypatch.delenv("LLAVA_ALIGN_REFERENCE_LLAVA", raising=False)
    _purge()
    saved = list(sys.path)
    sys.path.insert(0, ROOT)  # the repo's shims win over anything the driver appends
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    try:
        spec.loader.exec_module(mod)
    finally:
        sys.path[:] = saved
    return mod


class RecordingModel:
    """Stands where the engine-backed model would: every generate() call must bind to the real signature."""

    def __init__(self, vocab=512, layers=2, heads=2):
        from llava_align_b200.model import GenerateOutput, LlavaLlamaForCausalLM
        self._sig = inspect.signature(LlavaLlamaForCausalLM.generate)
        self._out = GenerateOutput
        self.config = argparse.Namespace(mm_use_im_start_end=False, mm_use_im_patch_token=False)
        self.calls, self.vocab, self.layers, self.heads = [], vocab, layers, heads

    def generate(self, *a, **kw):
        bound = self._sig.bind(self, *a, **kw)  # TypeError if a driver passes something generate() does not take
        self.calls.append(dict(bound.arguments))
        ids = bound.arguments["input_ids"]
        n = min(4, bound.arguments.get("max_new_tokens") or 4)
        seq = torch.cat([ids, torch.full((ids.shape[0], n), 7, dtype=ids.dtype)], dim=1)
        extra = bound.arguments.get("kwargs", {})
        if not extra.get("return_dict_in_generate"):
            return seq
        T = ids.shape[1]
        att = torch.softmax(torch.zeros(ids.shape[0], self.heads, T, T), -1)
        return self._out(sequences=seq, scores=tuple(torch.randn(ids.shape[0], self.vocab) for _ in range(n)),
                         attentions=tuple(tuple(att for _ in range(self.layers)) for _ in range(n)))


def _write_inputs(tmp_path):
    from PIL import Image
    img = Image.fromarray((np.random.default_rng(0).random((40, 64, 3)) * 255).astype(np.uint8))
    os.makedirs(tmp_path / "images", exist_ok=True)
    img.save(tmp_path / "images" / "a.png")
    qf = tmp_path / "POPE_q.jsonl"
    with open(qf, "w") as f:
        for i in range(2):
            f.write(json.dumps({"question_id": i, "image": "a.png", "text": "Is there a dog in the image?", "label": "yes"}) + "\n")
    return str(tmp_path / "images"), str(qf)


needs_ref = pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree exists only in the authoring container")


@needs_ref
@pytest.mark.parametrize("use_cd", [False, True])
def test_reference_llava_sampling_driver_runs_unchanged(tmp_path, monkeypatch, use_cd):
    drv = _load_driver(os.path.join(REF, "experiments/eval/sampling/llava_sampling.py"), "ref_llava_sampling", monkeypatch)
    import llava_align_b200.builder as B
    import llava_align_b200.model as M
    assert drv.load_pretrained_model is B.load_pretrained_model          # llava_sampling.py:14 resolved to this repo's loader
    assert drv.evolve_vcd_sampling is M.evolve_vcd_sampling and M.vdd_sampler_installed()   # :24-25 armed the VDD sampler
    assert inspect.getsourcefile(drv.tokenizer_image_token).startswith(REF)  # the reference's own mm_utils, untouched
    assert inspect.getsourcefile(drv.conv_templates["llava_v1"].__class__).startswith(REF)
    assert inspect.getsourcefile(drv.add_diffusion_noise).startswith(ROOT)
    model = RecordingModel()
    monkeypatch.setattr(drv, "load_pretrained_model", lambda *a, **k: (FakeTokenizer(), model, FakeProcessor(), 2048))
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    folder, qf = _write_inputs(tmp_path)
    args = argparse.Namespace(model_path="liuhaotian/llava-v1.5-7b", model_base=None, image_folder=folder, question_file=qf,
                              answers_file=str(tmp_path / "out" / "ans.jsonl"), conv_mode="llava_v1", num_chunks=1, chunk_idx=0,
                              temperature=1.0, top_p=1, top_k=None, noise_step=500, use_cd=use_cd, use_dd=True, use_dd_unk=True,
                              cd_alpha=1.0, cd_beta=0.1, seed=55)
    drv.eval_model(args)
    lines = [json.loads(l) for l in open(args.answers_file)]
    assert len(lines) == 2 and lines[0]["model_id"] == "llava-v1.5-7b"
    assert len(model.calls) == 2
    c = model.calls[0]
    assert (c["input_ids"] == -200).sum().item() == 1 and c["images"].dtype == torch.float16
    assert (c["images_cd"] is not None) == use_cd
    assert c["use_dd"] is True and c["use_dd_unk"] is True and c["cd_alpha"] == 1.0 and c["max_new_tokens"] == 1024


@needs_ref
def test_reference_llava_calibrate_driver_runs_unchanged(tmp_path, monkeypatch):
    monkeypatch.setattr(sys, "argv", ["llava_calibrate.py"])
    import types
    for dep in ("peft", "seaborn", "matplotlib", "matplotlib.pyplot"):  # plotting / LoRA imports the generate path never touches
        try:
            importlib.import_module(dep)
        except ImportError:
            stub = types.ModuleType(dep)
            stub.PeftModel = object
            monkeypatch.setitem(sys.modules, dep, stub)
    monkeypatch.delitem(sys.modules, "utils", raising=False)
    monkeypatch.delitem(sys.modules, "utils.metrics", raising=False)
    drv = _load_driver(os.path.join(REF, "experiments/eval/calibrate/llava_calibrate.py"), "ref_llava_calibrate", monkeypatch)
    import llava_align_b200.builder as B
    assert drv.load_pretrained_model is B.load_pretrained_model
    model = RecordingModel()
    monkeypatch.setattr(drv, "load_pretrained_model", lambda *a, **k: (FakeTokenizer(), model, FakeProcessor(), 2048))
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    folder, qf = _write_inputs(tmp_path)
    args = argparse.Namespace(model_path="liuhaotian/llava-v1.5-7b", model_base=None, image_folder=folder, question_file=qf,
                              answers_file=str(tmp_path / "out" / "ans.jsonl"), conv_mode="llava_v1", num_chunks=1, chunk_idx=0,
                              temperature=1.0, top_p=1, top_k=None, noise_step=500, use_cd=False, use_dd=True, use_dd_unk=True,
                              cd_alpha=1.0, cd_beta=0.1, seed=55)
    monkeypatch.setattr(drv, "args", args, raising=False)  # calibrate_label_sapce reads the module-level `args`
    try:
        drv.eval_model(args)
    except Exception as ex:  # the scorer tail of the driver (sklearn / file layout) is outside the generate path
        if not model.calls:
            raise
        tb = ex.__traceback__
        while tb.tb_next:
            tb = tb.tb_next
        assert "generate" not in tb.tb_frame.f_code.co_name, ex
    # per question: the VDD generate (:160-177) + the text-only and the <unk> generates (:192-194 -> :62-78)
    assert len(model.calls) >= 3
    main, none, unk = model.calls[:3]
    assert main["kwargs"]["output_attentions"] and main["kwargs"]["output_scores"] and main["kwargs"]["return_dict_in_generate"]
    assert main["max_new_tokens"] == 64 and main["use_dd"] is True
    assert none["images"] is None and (none["input_ids"] == -200).sum().item() == 0
    assert unk["images"] is None and (unk["input_ids"] == FakeTokenizer.unk_token_id).sum().item() >= 1


def test_fallback_host_helpers_match_reference_semantics(monkeypatch):
    """Without a reference tree the shim serves llava_align_b200.hostside; where the reference IS present its own functions
    must agree with the restatement on the driver's prompts."""
    from llava_align_b200 import hostside as H
    tok = FakeTokenizer()
    conv = H.conv_templates["llava_v1"].copy()
    conv.append_message(conv.roles[0], H.DEFAULT_IMAGE_TOKEN + "\nIs there a dog in the image? Please answer this question with one word.")
    conv.append_message(conv.roles[1], None)
    prompt = conv.get_prompt()
    assert prompt.endswith("ASSISTANT:") and prompt.count("<image>") == 1 and H.conv_templates["llava_v1"].messages == []
    ids = H.tokenizer_image_token(prompt, tok, H.IMAGE_TOKEN_INDEX, return_tensors="pt")
    assert ids[0] == tok.bos_token_id and (ids == -200).sum() == 1 and (ids == tok.bos_token_id).sum() == 1
    assert H.tokenizer_image_token("no image here", tok) == tok("no image here").input_ids
    assert H.tokenizer_image_token("<image><image> x", tok).count(-200) == 2
    assert H.get_model_name_from_path("/a/b/llava-v1.5-7b/") == "llava-v1.5-7b"
    assert H.get_model_name_from_path("/a/llava-x/checkpoint-300") == "llava-x_checkpoint-300"
    if not os.path.isdir(REF):
        return
    monkeypatch.setenv("LLAVA_ALIGN_REFERENCE_LLAVA", os.path.join(REF, "experiments", "llava"))
    _purge()
    saved = list(sys.path)
    sys.path.insert(0, ROOT)
    try:
        import llava.conversation as RC
        import llava.mm_utils as RM
    finally:
        sys.path[:] = saved
    rconv = RC.conv_templates["llava_v1"].copy()
    rconv.append_message(rconv.roles[0], H.DEFAULT_IMAGE_TOKEN + "\nIs there a dog in the image? Please answer this question with one word.")
    rconv.append_message(rconv.roles[1], None)
    assert rconv.get_prompt() == prompt and (rconv.sep, rconv.sep2) == (conv.sep, conv.sep2)
    for p in (prompt, "no image here", "<image><image> x", "a <image> b <image>", "<image>"):
        assert RM.tokenizer_image_token(p, tok) == H.tokenizer_image_token(p, tok), p
    assert RM.get_model_name_from_path("/a/llava-x/checkpoint-300") == H.get_model_name_from_path("/a/llava-x/checkpoint-300")
    from PIL import Image
    im = Image.fromarray((np.random.default_rng(1).random((30, 50, 3)) * 255).astype(np.uint8))
    bg = tuple(int(x * 255) for x in FakeProcessor.image_mean)
    assert np.array_equal(np.asarray(RM.expand2square(im, bg)), np.asarray(H.expand2square(im, bg)))
    assert np.array_equal(np.asarray(RM.expand2square(im.rotate(90, expand=True), bg)), np.asarray(H.expand2square(im.rotate(90, expand=True), bg)))
    _purge()


@needs_ref
def test_add_diffusion_noise_matches_reference():
    spec = importlib.util.spec_from_file_location("ref_add_noise", os.path.join(REF, "vcd_utils/vcd_add_noise.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    from vcd_utils.vcd_add_noise import add_diffusion_noise
    x = torch.randn(3, 24, 24)
    for step in (0, 1, 500, 999):
        torch.manual_seed(step)
        a = ref.add_diffusion_noise(x, step)
        torch.manual_seed(step)
        b = add_diffusion_noise(x, step)
        assert torch.equal(a, b), step


@pytest.mark.gpu
def test_driver_call_sites_on_the_engine(monkeypatch):
    """llava_sampling.py:96-114 and llava_calibrate.py:160-185,:62-82 statement for statement, on the tiny engine."""
    monkeypatch.delenv("LLAVA_ALIGN_REFERENCE_LLAVA", raising=False)
    _purge()
    sys.path.insert(0, ROOT)
    from llava.constants import IMAGE_TOKEN_INDEX, DEFAULT_IMAGE_TOKEN
    from llava.conversation import conv_templates, SeparatorStyle
    from llava.mm_utils import tokenizer_image_token, KeywordsStoppingCriteria
    from llava.model.builder import load_pretrained_model
    from vcd_utils.vcd_add_noise import add_diffusion_noise
    from vcd_utils.vcd_sample import evolve_vcd_sampling
    evolve_vcd_sampling()
    tokenizer = FakeTokenizer()
    _, model, _, context_len = load_pretrained_model("random:tiny:0", None, "tiny")
    cfg = model.vdd_config
    qs = DEFAULT_IMAGE_TOKEN + "\n" + "Is there a dog in the image?"
    conv = conv_templates["llava_v1"].copy()
    conv.append_message(conv.roles[0], qs + " Please answer this question with one word.")
    conv.append_message(conv.roles[1], None)
    prompt = conv.get_prompt()
    input_ids = tokenizer_image_token(prompt, tokenizer, IMAGE_TOKEN_INDEX, return_tensors="pt").unsqueeze(0).cuda()
    image_tensor = torch.randn(3, cfg.image_size, cfg.image_size)
    image_tensor_cd = add_diffusion_noise(image_tensor, 500)
    stop_str = conv.sep if conv.sep_style != SeparatorStyle.TWO else conv.sep2
    KeywordsStoppingCriteria([stop_str], tokenizer, input_ids)
    for cd in (None, image_tensor_cd):
        with torch.inference_mode():  # llava_sampling.py:95-109
            output_ids = model.generate(
                input_ids,
                images=image_tensor.unsqueeze(0).half().cuda(),
                images_cd=(cd.unsqueeze(0).half().cuda() if cd is not None else None),
                cd_alpha=1.0, cd_beta=0.1, use_dd=True, use_dd_unk=True,
                do_sample=True, temperature=1.0, top_p=1, top_k=None,
                max_new_tokens=12, use_cache=True)
        input_token_len = input_ids.shape[1]
        assert (input_ids != output_ids[:, :input_token_len]).sum().item() == 0   # llava_sampling.py:111-114
        tokenizer.batch_decode(output_ids[:, input_token_len:], skip_special_tokens=True)[0].strip()
    with torch.inference_mode():  # llava_calibrate.py:160-182
        model_outputs = model.generate(
            input_ids,
            images=image_tensor.unsqueeze(0).half().cuda(),
            images_cd=None, cd_alpha=1.0, cd_beta=0.1, use_dd=True, use_dd_unk=True,
            do_sample=True, temperature=1.0, top_p=1, top_k=None,
            max_new_tokens=8, use_cache=True, output_attentions=True, output_scores=True, return_dict_in_generate=True)
        output_ids = model_outputs['sequences']
        scores = model_outputs['scores'][0]
        attentions = model_outputs['attentions'][0][-1]
        attention = torch.mean(attentions, dim=1).squeeze()
    assert scores.shape == (1, cfg.vocab_size) and attention.shape[-1] == input_ids.shape[1] + cfg.num_patches - 1
    probs = torch.softmax(scores, dim=-1).float().cpu()   # utils/metrics.py:102-104
    assert abs(probs.sum().item() - 1) < 1e-4 and torch.topk(probs, k=10).values.shape == (1, 10)
    # llava_calibrate.py:192-194 -> :62-82: the text-only and the <unk> generates
    for variant in ("none", "unk"):
        q2 = (DEFAULT_IMAGE_TOKEN + "\n" if variant == "unk" else "") + "Is there a dog in the image?"
        c2 = conv_templates["llava_v1"].copy()
        c2.append_message(c2.roles[0], q2 + " Please answer this question with one word.")
        c2.append_message(c2.roles[1], None)
        ids2 = tokenizer_image_token(c2.get_prompt(), tokenizer, IMAGE_TOKEN_INDEX, return_tensors="pt").unsqueeze(0).cuda()
        if variant == "unk":
            ids2[ids2 == IMAGE_TOKEN_INDEX] = tokenizer.unk_token_id
        with torch.inference_mode():
            mo = model.generate(ids2, images=None, images_cd=None, cd_alpha=1.0, cd_beta=0.1, do_sample=True, temperature=1.0,
                                top_p=1, top_k=None, max_new_tokens=6, use_cache=True, output_scores=True, output_attentions=True,
                                return_dict_in_generate=True)
        assert mo['sequences'].shape[1] <= ids2.shape[1] + 6 and mo['scores'][0].shape == (1, cfg.vocab_size)
        assert mo['attentions'][0][-1].shape[-1] == ids2.shape[1]
    _purge()


def test_attention_capture_budget(monkeypatch):
    """output_attentions=True keeps the reference's whole `attentions` tuple (vcd_sample.py:241-246) while it fits the budget and
    falls back to the one entry the drivers read (`[0][-1]`, llava_calibrate.py:180) beyond it; the environment overrides both."""
    from llava_align_b200.model import _attention_capture_mode
    monkeypatch.delenv("VDD_ATTENTIONS", raising=False)
    monkeypatch.delenv("VDD_ATTENTIONS_MAX_GB", raising=False)
    # 7B, 607-token prompt after the image splice: 256 new tokens = 2.3 GB -> all; the drivers' max_new_tokens = 1024 = 8.3 GB -> last
    assert _attention_capture_mode(1, 32, 32, 607, 256) == "all"
    assert _attention_capture_mode(1, 32, 32, 607, 1024) == "last"
    assert _attention_capture_mode(32, 32, 32, 639, 64) == "last"      # a POPE batch of 32
    assert _attention_capture_mode(2, 2, 2, 27, 8) == "all"            # the tiny golden cases
    monkeypatch.setenv("VDD_ATTENTIONS_MAX_GB", "16")
    assert _attention_capture_mode(1, 32, 32, 607, 1024) == "all"
    monkeypatch.setenv("VDD_ATTENTIONS", "last")
    assert _attention_capture_mode(2, 2, 2, 27, 8) == "last"
    monkeypatch.setenv("VDD_ATTENTIONS", "all")
    assert _attention_capture_mode(32, 32, 32, 639, 1024) == "all"
