"""ctypes binding of libvdd.so (include/vdd.h).  No fallback: if the CUDA library is missing this raises.

The library is built in-tree by `__graft_entry__.build()` / `make -C llava_align_b200/csrc` into
llava_align_b200/libvdd.so so that it travels with the source tree.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvdd.so")


class VddError(RuntimeError):
    pass


class vdd_config(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("hidden_size", "intermediate_size", "num_layers", "num_heads", "vocab_size",
                                          "max_position")] + \
               [("rms_eps", C.c_float), ("rope_theta", C.c_float)] + \
               [(n, C.c_int32) for n in ("v_hidden", "v_intermediate", "v_layers", "v_layers_used", "v_heads",
                                          "image_size", "patch_size")] + \
               [("v_ln_eps", C.c_float), ("tp_size", C.c_int32), ("tp_rank", C.c_int32)]


class vdd_sample_params(C.Structure):
    _fields_ = [("cd_alpha", C.c_float), ("log_cd_beta", C.c_float), ("temperature", C.c_float), ("top_p", C.c_float),
                ("top_k", C.c_int32), ("use_contrast", C.c_int32), ("eos_token_id", C.c_int64),
                ("pad_token_id", C.c_int64)]


# name -> (restype, argtypes); every symbol include/vdd.h declares
_P = C.c_void_p
SIGNATURES = {
    "vdd_last_error": (C.c_char_p, []),
    "vdd_version": (C.c_int, []),
    "vdd_logit_combine_sample": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int64, C.POINTER(vdd_sample_params), _P, _P,
                                           _P, _P, _P]),
    "vdd_bench_combine": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int64, C.POINTER(vdd_sample_params), _P, _P, C.c_int,
                                    C.POINTER(C.c_float), _P]),
    "vdd_create": (C.c_int, [C.POINTER(vdd_config), C.POINTER(_P)]),
    "vdd_destroy": (None, [_P]),
    "vdd_load_tensor": (C.c_int, [_P, C.c_char_p, _P, C.POINTER(C.c_int64), C.c_int, _P]),
    "vdd_missing_tensors": (C.c_int, [_P, C.c_char_p, C.c_size_t]),
    "vdd_encode_images": (C.c_int, [_P, _P, C.c_int, _P]),
    "vdd_image_features": (_P, [_P]),
    "vdd_encode_images_cd": (C.c_int, [_P, _P, C.c_int, _P]),
    "vdd_prefill_cd": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, _P]),
    "vdd_prefill": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "vdd_set_attention_capture": (C.c_int, [_P, _P, C.c_int64]),
    "vdd_set_attention_capture_all": (C.c_int, [_P, _P, C.c_int64, _P, C.c_int64]),
    "vdd_set_sampler_rng": (C.c_int, [_P, C.c_uint64, C.c_uint64]),
    "vdd_sampler_rng_advance": (C.c_uint64, [_P]),
    "vdd_debug_philox_exponential": (C.c_int, [C.c_uint64, C.c_uint64, C.c_int64, _P, _P]),
    "vdd_debug_gemm_tile_coords": (C.c_int, [C.c_int, C.c_int, C.c_int, _P, _P]),
    "vdd_debug_prefill_attn": (C.c_int, [_P, C.c_int64, C.c_int64, _P, _P, C.c_int64, C.c_int64, _P, _P, _P, _P, _P, C.c_int64, C.c_int,
                                         C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, _P]),
    "vdd_set_prompt_mask": (C.c_int, [_P, _P, C.c_int, C.c_int]),
    "vdd_set_stream_scores_capture": (C.c_int, [_P, _P, C.c_int64]),
    "vdd_decode_step": (C.c_int, [_P, _P, _P]),
    "vdd_logits": (_P, [_P]),
    "vdd_num_streams": (C.c_int, [_P]),
    "vdd_generate": (C.c_int, [_P, _P, C.c_int, C.c_int, _P, C.c_int, C.c_int, C.POINTER(vdd_sample_params), C.c_int, _P,
                               _P, _P, C.POINTER(C.c_int), _P]),
    "vdd_generate_cd": (C.c_int, [_P, _P, C.c_int, C.c_int, _P, _P, C.c_int, C.c_int, C.POINTER(vdd_sample_params), C.c_int, _P,
                                  _P, _P, C.POINTER(C.c_int), _P]),
    "vdd_tp_unique_id": (C.c_int, [_P]),
    "vdd_tp_init": (C.c_int, [_P, _P]),
    "vdd_tp_comm_handle": (C.c_int, [_P, _P]),
    "vdd_tp_comm_open": (C.c_int, [_P, _P]),
    "vdd_launch_count": (C.c_int64, [_P, C.c_int]),
    "vdd_last_timings": (C.c_int, [_P, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "vdd_bench_projection_chain": (C.c_int, [_P, _P]),
    "vdd_debug_trace": (C.c_int, [_P, _P, C.c_int]),
    "vdd_gemv": (C.c_int, [_P, _P, _P, _P, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P]),
    "vdd_clip_preprocess": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P, _P, C.c_int, _P]),
    "vdd_gemm": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P, _P, C.c_int64, _P]),
}

_lib = None


def load():
    """Returns the loaded library; raises VddError when it has not been built (there is no CPU path)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise VddError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                       f"or `make -C llava_align_b200/csrc`.  There is no CPU fallback for the VDD path.")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().vdd_last_error().decode("utf-8", "replace")
        raise VddError(f"{what} failed ({rc}): {msg}")
