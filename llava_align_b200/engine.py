"""Python handle on a libvdd engine: weights in, generate() out.  PyTorch is used for device memory and streams only.

Mirrors what the reference does in Python around HF modules: from_pretrained (experiments/llava/model/builder.py:97-103),
the per-token loop of vcd_utils/vcd_sample.py:93-299 and LlavaLlamaForCausalLM.forward (llava_llama.py:58-128) -- all of
which run inside libvdd here.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional

import torch

from . import _lib
from .config import VddConfig, IMAGE_TOKEN_INDEX


class _DevPtr:
    """Exposes a raw device pointer to torch through __cuda_array_interface__ (no copy)."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 3,
                                         "strides": None}


def _stream_ptr(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def make_sample_params(cd_alpha=0.5, cd_beta=0.1, temperature=None, top_k=None, top_p=None, use_contrast=True,
                       eos_token_id=None, pad_token_id=0) -> _lib.vdd_sample_params:
    p = _lib.vdd_sample_params()
    p.cd_alpha = float(cd_alpha)
    # reference vcd_sample.py:191 evaluates torch.log(torch.tensor(cd_beta)) in fp32; use the very same op
    p.log_cd_beta = float(torch.log(torch.tensor(float(cd_beta))).item()) if cd_beta > 0 else -math.inf
    p.temperature = 1.0 if temperature is None else float(temperature)
    p.top_p = 1.0 if top_p is None else float(top_p)
    p.top_k = 0 if top_k is None else int(top_k)
    p.use_contrast = 1 if use_contrast else 0
    p.eos_token_id = -1 if eos_token_id is None else int(eos_token_id)
    p.pad_token_id = 0 if pad_token_id is None else int(pad_token_id)
    return p


def logit_combine_sample(l, u=None, d=None, cd_alpha=0.5, cd_beta=0.1, temperature=None, top_k=None, top_p=None,
                         q=None, want_scores=False, unfinished=None, eos_token_id=None, pad_token_id=0):
    """One launch of K8 on fp32 CUDA tensors [B,V]; returns (tokens int64 [B], scores or None)."""
    lib = _lib.load()
    assert l.is_cuda and l.dtype == torch.float32 and l.dim() == 2 and l.stride(1) == 1
    B, V = l.shape
    for t in (u, d, q):
        assert t is None or (t.is_cuda and t.dtype == torch.float32 and t.shape == l.shape and t.stride() == l.stride())
    if q is not None:
        assert q.is_contiguous()
    tokens = torch.empty(B, dtype=torch.int64, device=l.device)
    scores = torch.empty(B, V, dtype=torch.float32, device=l.device) if want_scores else None
    p = make_sample_params(cd_alpha, cd_beta, temperature, top_k, top_p, u is not None or d is not None, eos_token_id,
                           pad_token_id)
    ptr = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
    _lib.check(lib.vdd_logit_combine_sample(ptr(l), ptr(u), ptr(d), B, V, l.stride(0), C.byref(p), ptr(q), ptr(scores),
                                            ptr(tokens), ptr(unfinished), _stream_ptr(l.device)), "vdd_logit_combine_sample")
    return tokens, scores


# how each HF tensor is sliced for tensor parallelism: (dim, which local size)
def _tp_slice(name: str, t: torch.Tensor, tp: int, rank: int) -> torch.Tensor:
    if tp == 1:
        return t
    def rows(x):
        n = x.shape[0] // tp
        return x[rank * n:(rank + 1) * n]
    def cols(x):
        n = x.shape[1] // tp
        return x[:, rank * n:(rank + 1) * n]
    if name.startswith("model.layers."):
        if any(name.endswith(s) for s in ("q_proj.weight", "k_proj.weight", "v_proj.weight", "gate_proj.weight",
                                          "up_proj.weight")):
            return rows(t)
        if name.endswith("o_proj.weight") or name.endswith("down_proj.weight"):
            return cols(t)
    if name == "lm_head.weight":
        return rows(t)
    return t


class VddEngine:
    def __init__(self, cfg: VddConfig, device="cuda:0", tp_size: int = 1, tp_rank: int = 0):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.VddError("the VDD engine needs a CUDA device (sm_100a); there is no CPU path")
        self.cfg = cfg
        self.device = torch.device(device)
        self.tp_size, self.tp_rank = tp_size, tp_rank
        torch.cuda.set_device(self.device)
        c = _lib.vdd_config()
        c.hidden_size, c.intermediate_size, c.num_layers = cfg.hidden_size, cfg.intermediate_size, cfg.num_hidden_layers
        c.num_heads, c.vocab_size, c.max_position = cfg.num_attention_heads, cfg.vocab_size, cfg.max_position_embeddings
        c.rms_eps, c.rope_theta = cfg.rms_norm_eps, cfg.rope_theta
        c.v_hidden, c.v_intermediate, c.v_layers = cfg.vision_hidden_size, cfg.vision_intermediate_size, cfg.vision_num_hidden_layers
        c.v_layers_used, c.v_heads = cfg.vision_layers_used, cfg.vision_num_attention_heads
        c.image_size, c.patch_size, c.v_ln_eps = cfg.image_size, cfg.patch_size, cfg.vision_layer_norm_eps
        c.tp_size, c.tp_rank = tp_size, tp_rank
        h = C.c_void_p()
        _lib.check(self.lib.vdd_create(C.byref(c), C.byref(h)), "vdd_create")
        self.h = h
        self._keep = []  # tensors whose memory the library is reading asynchronously
        self.stream = torch.cuda.Stream(self.device)  # graph capture needs a non-legacy stream

    def close(self):
        if getattr(self, "h", None):
            self.lib.vdd_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- weights
    def load_tensor(self, name: str, t: torch.Tensor):
        t = _tp_slice(name, t, self.tp_size, self.tp_rank)
        t = t.to(self.device, torch.bfloat16).contiguous()
        shape = (C.c_int64 * t.dim())(*t.shape)
        _lib.check(self.lib.vdd_load_tensor(self.h, name.encode(), C.c_void_p(t.data_ptr()), shape, t.dim(),
                                            _stream_ptr(self.device)), f"vdd_load_tensor({name})")
        torch.cuda.current_stream(self.device).synchronize()  # t may be freed by the caller right after

    def load_state_dict(self, sd: dict):
        for k, v in sd.items():
            self.load_tensor(k, v)
        self.check_complete()

    def check_complete(self):
        buf = C.create_string_buffer(1024)
        n = self.lib.vdd_missing_tensors(self.h, buf, 1024)
        if n:
            raise _lib.VddError(f"{n} tensors missing: {buf.value.decode()}")

    def tp_init(self, id_bytes: bytes):
        b = C.create_string_buffer(id_bytes, 128)
        _lib.check(self.lib.vdd_tp_init(self.h, b), "vdd_tp_init")

    def tp_setup(self, dist):
        """Join the tensor-parallel group through an initialised torch.distributed process group: NCCL communicator for
        the GEMM-sized paths, cudaIpc-mapped peer buffers for the in-kernel all-reduce of the decode chain."""
        dev = self.device
        idt = torch.zeros(128, dtype=torch.uint8, device=dev)
        if self.tp_rank == 0:
            idt = torch.frombuffer(bytearray(self.tp_unique_id()), dtype=torch.uint8).to(dev)
        dist.broadcast(idt, 0)
        self.tp_init(bytes(idt.cpu().numpy().tobytes()))
        h = C.create_string_buffer(64)
        _lib.check(self.lib.vdd_tp_comm_handle(self.h, h), "vdd_tp_comm_handle")
        mine = torch.frombuffer(bytearray(h.raw), dtype=torch.uint8).to(dev)
        allh = [torch.zeros(64, dtype=torch.uint8, device=dev) for _ in range(self.tp_size)]
        dist.all_gather(allh, mine)
        blob = b"".join(bytes(t.cpu().numpy().tobytes()) for t in allh)
        _lib.check(self.lib.vdd_tp_comm_open(self.h, C.create_string_buffer(blob, len(blob))), "vdd_tp_comm_open")
        dist.barrier()
        self._dist = dist

    def tp_broadcast(self, t: torch.Tensor) -> torch.Tensor:
        """Make a per-rank tensor (the sampler's Exp(1) variates) identical on every tensor-parallel rank: every rank samples
        the same token from the same all-gathered logits only if it races against the same variates, and ranks that pick
        different tokens would silently diverge in their sharded KV caches."""
        if self.tp_size > 1:
            if getattr(self, "_dist", None) is None:
                raise _lib.VddError("tensor-parallel engine without tp_setup(): cannot synchronise the sampler's variates")
            self._dist.broadcast(t, 0)
        return t

    @staticmethod
    def tp_unique_id() -> bytes:
        b = C.create_string_buffer(128)
        _lib.check(_lib.load().vdd_tp_unique_id(b), "vdd_tp_unique_id")
        return b.raw

    # ---- step-wise API (parity tests, host-driven sampling)
    def encode_images(self, images: torch.Tensor) -> torch.Tensor:
        px = images.to(self.device, torch.bfloat16).contiguous()
        B = px.shape[0]
        _lib.check(self.lib.vdd_encode_images(self.h, C.c_void_p(px.data_ptr()), B, _stream_ptr(self.device)), "vdd_encode_images")
        self._keep = [px]
        P, H = self.cfg.num_patches, self.cfg.hidden_size
        return torch.as_tensor(_DevPtr(self.lib.vdd_image_features(self.h), (B, P, H), "<f4"), device=self.device)

    def set_sampler_rng(self, seed: int, offset: int):
        """The next generate() without an explicit q draws its Exp(1) variates inside K8 from this torch-generator state."""
        _lib.check(self.lib.vdd_set_sampler_rng(self.h, C.c_uint64(seed & (2 ** 64 - 1)), C.c_uint64(offset)), "vdd_set_sampler_rng")

    def sampler_rng_advance(self) -> int:
        return int(self.lib.vdd_sampler_rng_advance(self.h))

    def set_prompt_mask(self, attention_mask: Optional[torch.Tensor]):
        """Ragged batch: [B, T] mask (nonzero = token) for the next prefill / generate; None disarms."""
        if attention_mask is None:
            _lib.check(self.lib.vdd_set_prompt_mask(self.h, None, 0, 0), "vdd_set_prompt_mask")
            return
        m = (attention_mask.detach().to("cpu") != 0).to(torch.uint8).contiguous()
        _lib.check(self.lib.vdd_set_prompt_mask(self.h, C.c_void_p(m.data_ptr()), m.shape[0], m.shape[1]), "vdd_set_prompt_mask")

    def prefill(self, input_ids: torch.Tensor, with_images: bool, use_unk: bool, use_none: bool, max_new_tokens: int,
                attention_mask: Optional[torch.Tensor] = None):
        ids = input_ids.detach().to("cpu", torch.int64).contiguous()
        B, T = ids.shape
        if attention_mask is not None:
            self.set_prompt_mask(attention_mask)
        _lib.check(self.lib.vdd_prefill(self.h, C.c_void_p(ids.data_ptr()), B, T, int(with_images), int(use_unk), int(use_none),
                                        int(max_new_tokens), _stream_ptr(self.device)), "vdd_prefill")
        self._B = B
        return self.logits()

    def decode_step(self, tokens: torch.Tensor):
        tk = tokens.to(self.device, torch.int64).contiguous()
        _lib.check(self.lib.vdd_decode_step(self.h, C.c_void_p(tk.data_ptr()), _stream_ptr(self.device)), "vdd_decode_step")
        self._keep = [tk]
        return self.logits()

    def logits(self) -> torch.Tensor:
        S = self.lib.vdd_num_streams(self.h)
        return torch.as_tensor(_DevPtr(self.lib.vdd_logits(self.h), (S, self._B, self.cfg.vocab_size), "<f4"), device=self.device)

    # ---- fused generate (the product path): host buffers in, host tokens out
    def generate(self, input_ids: torch.Tensor, images: Optional[torch.Tensor], use_unk: bool, use_none: bool,
                 params: _lib.vdd_sample_params, max_new_tokens: int, q: Optional[torch.Tensor] = None,
                 scores_out: Optional[torch.Tensor] = None, images_cd: Optional[torch.Tensor] = None,
                 attention_out: Optional[torch.Tensor] = None, stream_scores_out: Optional[torch.Tensor] = None,
                 attention_mask: Optional[torch.Tensor] = None, attention_all_out=None):
        """attention_out: optional fp32 [B, heads, len, len] device tensor that receives the step-0 attention map of the
        last layer of the main stream (vdd_set_attention_capture).  attention_all_out: optional pair of fp32 device tensors
        ([L, B, heads, len, len], [max_new-1, L, B, heads, len+max_new-1] or None) for every layer and step
        (vdd_set_attention_capture_all)."""
        ids = input_ids.detach().to("cpu", torch.int64).contiguous()
        B, T = ids.shape
        # the library copies B*3*S*S elements from the raw pointer: refuse anything else (the reference also accepts lists and
        # 5-D multi-crop tensors, llava_arch.py:97 -- outside the one-image-per-row path)
        for what, t in (("images", images), ("images_cd", images_cd)):
            if t is None:
                continue
            S = self.cfg.image_size
            if not isinstance(t, torch.Tensor) or tuple(t.shape) != (B, 3, S, S):
                got = tuple(t.shape) if isinstance(t, torch.Tensor) else type(t).__name__
                raise ValueError(f"{what} must be a tensor of shape {(B, 3, S, S)} (one {S}px image per prompt row), got {got}")
        if attention_mask is not None:
            self.set_prompt_mask(attention_mask)
        if stream_scores_out is not None:
            _lib.check(self.lib.vdd_set_stream_scores_capture(self.h, C.c_void_p(stream_scores_out.data_ptr()), stream_scores_out.numel()),
                       "vdd_set_stream_scores_capture")
        if attention_out is not None:
            _lib.check(self.lib.vdd_set_attention_capture(self.h, C.c_void_p(attention_out.data_ptr()), attention_out.numel()),
                       "vdd_set_attention_capture")
        if attention_all_out is not None:
            a0, an = attention_all_out
            _lib.check(self.lib.vdd_set_attention_capture_all(self.h, C.c_void_p(a0.data_ptr()), a0.numel(),
                                                              C.c_void_p(an.data_ptr()) if an is not None else None,
                                                              an.numel() if an is not None else 0), "vdd_set_attention_capture_all")
        px = None
        if images is not None:
            # host (ideally pinned) pixels are copied by the library inside the call; device pixels are used in place
            px = images.detach().to(torch.bfloat16).contiguous()
        pxc = None
        if images_cd is not None:
            pxc = images_cd.detach().to(torch.bfloat16).contiguous()
        out = torch.empty(B, max_new_tokens, dtype=torch.int64).pin_memory()
        steps = C.c_int(0)
        ptr = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
        self.stream.wait_stream(torch.cuda.current_stream(self.device))  # q / scores_out were produced there
        _lib.check(self.lib.vdd_generate_cd(self.h, ptr(ids), B, T, ptr(px), ptr(pxc), int(use_unk), int(use_none),
                                            C.byref(params), int(max_new_tokens), ptr(q), ptr(scores_out), ptr(out),
                                            C.byref(steps), C.c_void_p(self.stream.cuda_stream)), "vdd_generate")  # synchronised
        self._B = B
        return out[:, :steps.value], steps.value

    def launch_count(self, reset=False) -> int:
        return int(self.lib.vdd_launch_count(self.h, int(reset)))

    def last_timings(self):
        v, p, d = C.c_float(), C.c_float(), C.c_float()
        self.lib.vdd_last_timings(self.h, C.byref(v), C.byref(p), C.byref(d))
        return {"vision_ms": v.value, "prefill_ms": p.value, "decode_ms": d.value}
