"""CPU-side checks of the drop-in boundary: libvdd.so loads, exports every symbol include/vdd.h declares, its ctypes
structs match the header, and the product path refuses to run without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header():
    return open(os.path.join(ROOT, "include", "vdd.h")).read()


def test_library_exports_every_declared_symbol():
    from llava_align_b200 import _lib
    declared = set(re.findall(r"VDD_API\s+[\w\s\*]+?\b(vdd_\w+)\s*\(", _header()))
    assert len(declared) >= 20
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), f"libvdd.so does not export {name}"
    # and the Python binding declares a signature for each of them
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert lib.vdd_version() >= 100


def test_struct_layouts_match_header():
    from llava_align_b200 import _lib
    h = _header()
    cfg_body = re.search(r"typedef struct vdd_config \{(.*?)\} vdd_config;", h, re.S).group(1)
    fields = []
    for decl in re.findall(r"(int32_t|float)\s+([^;]+);", cfg_body):
        fields += [(n.strip(), decl[0]) for n in decl[1].split(",")]
    got = [(n, "float" if t is C.c_float else "int32_t") for n, t in _lib.vdd_config._fields_]
    assert got == fields
    sp_body = re.search(r"typedef struct vdd_sample_params \{(.*?)\} vdd_sample_params;", h, re.S).group(1)
    sp = [(n, t) for t, n in re.findall(r"(int32_t|int64_t|float)\s+(\w+);", sp_body)]
    names = {C.c_float: "float", C.c_int32: "int32_t", C.c_int64: "int64_t"}
    assert [(n, names[t]) for n, t in _lib.vdd_sample_params._fields_] == sp
    assert C.sizeof(_lib.vdd_sample_params) == 40


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    from llava_align_b200 import _lib
    from llava_align_b200.config import tiny_config
    from llava_align_b200.model import LlavaLlamaForCausalLM
    with pytest.raises(_lib.VddError):
        LlavaLlamaForCausalLM(tiny_config(), "cuda:0")
    lib = _lib.load()
    h = C.c_void_p()
    c = _lib.vdd_config()
    assert lib.vdd_create(C.byref(c), C.byref(h)) != 0
    assert b"CUDA device" in lib.vdd_last_error()


def test_product_path_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "llava_align_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f"{fn} imports the oracle"
            assert "vdd_oracle" not in src and "cpu_baseline" not in src, fn
    assert not re.search(r"^\s*(from|import)\s+oracle", open(os.path.join(ROOT, "vcd_utils", "vcd_sample.py")).read(), re.M)


def test_generate_kwarg_surface():
    """The kwargs the reference drivers pass (llava_sampling.py:96-109, llava_calibrate.py:160-177) are all accepted."""
    import inspect
    from llava_align_b200.model import LlavaLlamaForCausalLM, evolve_vcd_sampling, vdd_sampler_installed
    sig = inspect.signature(LlavaLlamaForCausalLM.generate)
    for k in ("input_ids", "images", "images_cd", "cd_alpha", "cd_beta", "use_dd", "use_dd_unk", "max_new_tokens", "use_cache",
              "num_beams"):
        assert k in sig.parameters
    evolve_vcd_sampling()
    evolve_vcd_sampling()  # idempotent
    assert vdd_sampler_installed()
    from vcd_utils.vcd_sample import evolve_vcd_sampling as e2  # the reference's import path
    assert e2 is evolve_vcd_sampling


def test_null_arguments_return_error_codes_not_crashes():
    """Error behaviour of the C ABI (include/vdd.h: 0 on success, negative code + vdd_last_error() otherwise), exercised on
    the argument checks that run before any CUDA call."""
    from llava_align_b200 import _lib
    lib = _lib.load()
    assert lib.vdd_set_attention_capture(None, None, 0) < 0 and lib.vdd_last_error()
    assert lib.vdd_decode_step(None, None, None) < 0
    assert lib.vdd_prefill(None, None, 1, 1, 1, 0, 0, 1, None) < 0
    assert lib.vdd_bench_projection_chain(None, None) < 0
    assert lib.vdd_load_tensor(None, b"x", None, None, 0, None) < 0
    assert lib.vdd_num_streams(None) == 0 and lib.vdd_logits(None) is None
    assert lib.vdd_launch_count(None, 0) == 0


def test_attentions_tuple_serves_only_what_is_materialised():
    """`attentions[0][-1]` is the one entry the reference's drivers read (llava_calibrate.py:180); anything else must raise
    rather than return something made up."""
    from llava_align_b200.model import _StepAttentions
    m = torch.zeros(1, 2, 5, 5)
    a0, a1 = _StepAttentions(0, 4, m), _StepAttentions(1, 4, None)
    assert len(a0) == 4 and a0[-1] is m and a0[3] is m
    for bad in (lambda: a0[0], lambda: a0[2], lambda: a1[-1], lambda: list(a0)):
        with pytest.raises(NotImplementedError):
            bad()


@pytest.mark.parametrize("m_tiles,n_tiles", [(3, 48), (96, 48), (96, 86), (20, 4), (1, 86), (17, 1), (16, 5), (33, 7)])
def test_gemm_tile_walk_is_a_bijection_that_shares_panels(m_tiles, n_tiles):
    """The persistent tcgen05 GEMMs walk their tiles in groups of 16 m-tiles x all n-tiles (gemm_tc.cu, gemm_tile_coords): every
    (m, n) tile exactly once, and the 74 tiles in flight at any moment (one per SM pair) touch few distinct operand panels -- with
    the plain m-fastest walk they touched 74 A panels, which at 24 512 prefill rows made the GEMM DRAM-bound.  Host function."""
    from llava_align_b200 import _lib
    lib = _lib.load()
    seen, order = set(), []
    m, n = C.c_int(-1), C.c_int(-1)
    for t in range(m_tiles * n_tiles):
        assert lib.vdd_debug_gemm_tile_coords(t, m_tiles, n_tiles, C.byref(m), C.byref(n)) == 0
        assert 0 <= m.value < m_tiles and 0 <= n.value < n_tiles
        seen.add((m.value, n.value))
        order.append((m.value, n.value))
    assert len(seen) == m_tiles * n_tiles
    assert lib.vdd_debug_gemm_tile_coords(m_tiles * n_tiles, m_tiles, n_tiles, C.byref(m), C.byref(n)) != 0  # out of range
    wave = 74
    for t0 in range(0, len(order) - wave + 1, wave):
        w = order[t0:t0 + wave]
        panels = len({a for a, _ in w}) + len({b for _, b in w})
        # a window inside one group of gm m-tiles spans <= gm A panels and ceil(74 / gm) + 1 B panels; one crossing a group boundary
        # at most twice that (m-fastest over all of M would be min(74, m_tiles) A panels)
        gm = min(16, m_tiles)
        assert panels <= 2 * (gm + -(-wave // gm) + 1), (t0, panels)
