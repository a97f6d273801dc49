"""GPU parity of the stand-alone kernels, called through the C ABI (ctypes), against the CPU oracle / torch fp32."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import vdd_oracle as O


def _lib():
    from llava_align_b200 import _lib
    return _lib


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


# ------------------------------------------------------------------------------------------------ K8
def _logit_case(B, V, seed, spread=2.5):
    g = torch.Generator().manual_seed(seed)
    l = torch.randn(B, V, generator=g) * spread
    u = l + torch.randn(B, V, generator=g) * 0.7
    d = l + torch.randn(B, V, generator=g) * 0.7
    q = torch.empty(B, V).exponential_(1, generator=g)
    return l, u, d, q


CASES = [
    # alpha, beta, T, top_k, top_p, streams
    (1.0, 0.1, None, 1, None, "ud"),
    (1.0, 0.1, None, None, None, "ud"),
    (0.5, 0.1, None, None, None, "u"),
    (1.0, 0.2, 0.7, 5, None, "d"),
    (0.5, 0.01, None, None, 0.9, "u"),
    (1.0, 0.1, 0.3, 50, 0.8, "ud"),
    (1.0, 1e-4, None, 200, 0.95, "ud"),   # many survivors -> large sort
    (1.0, 0.1, None, None, 0.0, "ud"),    # top_p = 0 keeps the single best (llava_sampling.py:176 sweeps from 0)
    (2.0, 0.5, 1.5, 2, 0.5, "u"),
    (0.0, 0.1, None, None, None, ""),     # plain sampling, no contrast stream
    (0.0, 0.1, 0.9, 40, 0.6, ""),         # plain sampling of all V entries through the sort path
]


# B < 96: one 8-CTA cluster (or one sorting CTA) per row;  B >= 96 without top-k / top-p ordering: the row-streaming persistent kernel
@pytest.mark.parametrize("V,B", [(32000, 4), (512, 3), (1000, 2), (32000, 9), (1000, 12), (32000, 200)])
@pytest.mark.parametrize("case", CASES)
def test_combine_sample_matches_oracle(case, V, B):
    from llava_align_b200.engine import logit_combine_sample
    alpha, beta, T, k, p, streams = case
    l, u, d, q = _logit_case(B, V, seed=hash((V, B)) % 1000 + len(streams))
    uu = u if "u" in streams else None
    dd = d if "d" in streams else None
    tok_ref, s_ref = O.vdd_step(l, uu, dd, alpha, beta, T, k, p, q)
    cu = lambda t: t.cuda() if t is not None else None
    tok, s = logit_combine_sample(cu(l), cu(uu), cu(dd), alpha, beta, T, k, p, q=cu(q), want_scores=True)
    s = s.cpu()
    # same plausibility / top-k / top-p sets and bit-identical fp32 scores
    assert torch.equal(torch.isinf(s), torch.isinf(s_ref)), (torch.isinf(s).sum(1), torch.isinf(s_ref).sum(1))
    fin = torch.isfinite(s_ref)
    assert torch.equal(s[fin], s_ref[fin])
    assert torch.equal(tok.cpu(), tok_ref)


def test_combine_sample_argmax_and_eos():
    from llava_align_b200.engine import logit_combine_sample
    B, V = 5, 32000
    l, u, d, _ = _logit_case(B, V, 7)
    tok_ref, _ = O.vdd_step(l, u, d, 1.0, 0.1, None, 1, None, None if False else torch.ones(B, V))
    unf = torch.tensor([1, 0, 1, 1, 0], dtype=torch.int32, device="cuda")
    eos = int(tok_ref[2])
    tok, _ = logit_combine_sample(l.cuda(), u.cuda(), d.cuda(), 1.0, 0.1, None, 1, None, q=None, unfinished=unf,
                                  eos_token_id=eos, pad_token_id=0)
    exp = tok_ref.clone()
    exp[1] = 0
    exp[4] = 0
    assert torch.equal(tok.cpu(), exp)
    exp_unf = torch.tensor([1, 0, 0, 1, 0], dtype=torch.int32)
    exp_unf[0] = int(exp[0] != eos)
    exp_unf[3] = int(exp[3] != eos)
    assert torch.equal(unf.cpu(), exp_unf)


def test_combine_sample_rows_kernel_eos_and_ties():
    """Row-streaming kernel (B >= 96): EOS / pad bookkeeping, exact ties under top_k = 1 resolved by the race, no-q arg-max."""
    from llava_align_b200.engine import logit_combine_sample
    B, V = 128, 32000
    l, u, d, q = _logit_case(B, V, 11)
    l[3, 100] = l[3, 2000] = l[3].max() + 1.0  # an exact tie of the raw maximum; u = d = l there so the combined scores tie too
    u[3, 100] = u[3, 2000] = d[3, 100] = d[3, 2000] = l[3, 100]
    for qq in (q, None):
        tok_ref, s_ref = O.vdd_step(l, u, d, 1.0, 0.1, None, 1, None, qq if qq is not None else torch.ones(B, V))
        unf = torch.ones(B, dtype=torch.int32, device="cuda")
        unf[[1, 7]] = 0
        eos = int(tok_ref[2])
        tok, s = logit_combine_sample(l.cuda(), u.cuda(), d.cuda(), 1.0, 0.1, None, 1, None, q=qq.cuda() if qq is not None else None,
                                      want_scores=True, unfinished=unf, eos_token_id=eos, pad_token_id=0)
        exp = tok_ref.clone()
        exp[[1, 7]] = 0
        assert torch.equal(tok.cpu(), exp)
        assert torch.equal(torch.isinf(s.cpu()), torch.isinf(s_ref)) and int(torch.isfinite(s_ref[3]).sum()) == 2
        exp_unf = torch.tensor([int(i not in (1, 7) and int(exp[i]) != eos) for i in range(B)], dtype=torch.int32)
        assert torch.equal(unf.cpu(), exp_unf)


@pytest.mark.parametrize("B,V", [(1, 32000), (2, 512), (3, 1000), (32, 32000), (64, 32000), (200, 32000)])
def test_in_kernel_variates_are_torchs_exponential_stream(B, V):
    """K8 can draw its Exp(1) variates itself from a torch generator state; they must be, bit for bit, what
    `torch.empty(B, V, device='cuda').exponential_(1)` -- the draw inside torch.multinomial(probs, 1), reference vcd_sample.py:202 --
    produces under that state: Philox4x32-10, ATen's thread / call / component mapping (grid capped at #SM * 8 blocks: B * V beyond
    303104 elements wraps onto the components, beyond 1.2 M onto further calls), -log(u), and the per-call offset stride."""
    L = _lib()
    lib = L.load()
    gen = torch.cuda.default_generators[torch.cuda.current_device()]
    torch.manual_seed(1234 + B)
    torch.rand(7, device="cuda")                      # move the offset off zero
    for call in range(2):
        seed, off = gen.initial_seed(), gen.get_offset()
        ref = torch.empty(B, V, device="cuda").exponential_(1)
        stride = gen.get_offset() - off
        got = torch.empty(B, V, device="cuda")
        L.check(lib.vdd_debug_philox_exponential(C.c_uint64(seed), C.c_uint64(off), B * V, _ptr(got), _stream()), "philox")
        assert torch.equal(got, ref), (call, (got != ref).sum().item(), got.flatten()[:4], ref.flatten()[:4])
        assert stride % 4 == 0 and stride > 0


def test_combine_sample_ties_split_by_top_p():
    """Equal scores straddling the top-p boundary: HF's stable ascending sort removes the LOWER indices first."""
    from llava_align_b200.engine import logit_combine_sample
    V = 512
    l = torch.full((1, V), -20.0)
    l[0, [3, 77, 200, 411]] = 1.0   # four tied candidates, p = 0.25 each
    l[0, 9] = 3.0
    q = torch.ones(1, V)
    for top_p in (0.3, 0.75, 0.9, 0.95):
        tok_ref, s_ref = O.vdd_step(l, None, None, 0.0, 0.1, None, None, top_p, q)
        tok, s = logit_combine_sample(l.cuda(), None, None, 0.0, 0.1, None, None, top_p, q=q.cuda(), want_scores=True)
        assert torch.equal(torch.isinf(s.cpu()), torch.isinf(s_ref)), top_p
        assert torch.equal(tok.cpu(), tok_ref)


# ------------------------------------------------------------------------------------------------ GEMV
@pytest.mark.parametrize("R", [1, 2, 3, 4, 6, 8])
@pytest.mark.parametrize("N,K", [(4096, 4096), (1024, 11008), (512, 1376), (32000, 256)])
def test_gemv_plain_and_add(R, N, K):
    L = _lib()
    lib = L.load()
    g = torch.Generator(device="cuda").manual_seed(R * 131 + N + K)
    W = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).bfloat16()
    x = torch.randn(R, K, device="cuda", generator=g).bfloat16()
    ref = x.float() @ W.float().T
    out = torch.zeros(R, N, device="cuda")
    L.check(lib.vdd_gemv(_ptr(W), _ptr(x), None, None, 0.0, R, N, K, 0, _ptr(out), None, _stream()), "gemv")
    torch.testing.assert_close(out, ref, rtol=2e-4, atol=2e-4)
    base = torch.randn(R, N, device="cuda", generator=g)
    out2 = base.clone()
    L.check(lib.vdd_gemv(_ptr(W), _ptr(x), None, None, 0.0, R, N, K, 1, _ptr(out2), None, _stream()), "gemv add")
    torch.testing.assert_close(out2, base + ref, rtol=2e-4, atol=2e-4)


@pytest.mark.parametrize("R", [1, 3, 8])
def test_gemv_rmsnorm_swiglu(R):
    L = _lib()
    lib = L.load()
    K, I = 4096, 2752
    g = torch.Generator(device="cuda").manual_seed(R)
    gate = (torch.randn(I, K, device="cuda", generator=g) / K ** 0.5).bfloat16()
    up = (torch.randn(I, K, device="cuda", generator=g) / K ** 0.5).bfloat16()
    W = torch.stack([gate, up], dim=1).reshape(2 * I, K).contiguous()  # interleaved rows
    resid = torch.randn(R, K, device="cuda", generator=g) * 3
    nw = (1 + 0.1 * torch.randn(K, device="cuda", generator=g)).bfloat16()
    eps = 1e-5
    xn = (resid * torch.rsqrt(resid.pow(2).mean(-1, keepdim=True) + eps) * nw.float()).bfloat16().float()
    gg, uu = xn @ gate.float().T, xn @ up.float().T
    ref = torch.nn.functional.silu(gg) * uu
    out = torch.zeros(R, I, device="cuda", dtype=torch.bfloat16)
    L.check(lib.vdd_gemv(_ptr(W), None, _ptr(resid), _ptr(nw), eps, R, 2 * I, K, 2, None, _ptr(out), _stream()), "gemv swiglu")
    torch.testing.assert_close(out.float(), ref, rtol=1e-2, atol=1e-2)
    # fp32 epilogue with the same prologue
    outf = torch.zeros(R, 2 * I, device="cuda")
    L.check(lib.vdd_gemv(_ptr(W), None, _ptr(resid), _ptr(nw), eps, R, 2 * I, K, 0, _ptr(outf), None, _stream()), "gemv norm")
    torch.testing.assert_close(outf, xn @ W.float().T, rtol=3e-3, atol=3e-3)


# ------------------------------------------------------------------------------------------------ tcgen05 GEMM
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (670, 768, 256), (300, 384, 1376), (1154, 1024, 640), (96, 4096, 4096),
                                   (5, 512, 128),
                                   # at least one wave of 256x256 tiles over the SM pairs -> the persistent cta_group::2 variant (TMEM
                                   # double buffer), incl. N off the 256 tile and several tiles per pair
                                   (1300, 4096, 512), (2000, 3088, 320), (670, 12288, 4096),
                                   # cta_group::2 pair kernel (>= 74 tiles of 256x256): more m-tiles than one raster group (16) with a
                                   # short last group, and M between one and two CTA row blocks (the second CTA of the pair mostly empty)
                                   (5000, 1024, 256), (200, 22016, 512)])
def test_gemm_epilogues(M, N, K):
    L = _lib()
    lib = L.load()
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    W = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).bfloat16()
    bias = (0.1 * torch.randn(N, device="cuda", generator=g)).bfloat16()
    ref = A.float() @ W.float().T
    tol = dict(rtol=2e-3, atol=2e-3)
    out = torch.full((M, N), 7.0, device="cuda")
    L.check(lib.vdd_gemm(_ptr(A), _ptr(W), M, N, K, 0, None, None, _ptr(out), None, N, _stream()), "gemm f32")
    torch.testing.assert_close(out, ref, **tol)
    base = torch.randn(M, N, device="cuda", generator=g)
    out = base.clone()
    L.check(lib.vdd_gemm(_ptr(A), _ptr(W), M, N, K, 1, _ptr(bias), None, _ptr(out), None, N, _stream()), "gemm add")
    torch.testing.assert_close(out, base + ref + bias.float(), **tol)
    ob = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
    L.check(lib.vdd_gemm(_ptr(A), _ptr(W), M, N, K, 2, _ptr(bias), None, None, _ptr(ob), N, _stream()), "gemm bf16")
    torch.testing.assert_close(ob.float(), ref + bias.float(), rtol=1e-2, atol=1e-2)
    L.check(lib.vdd_gemm(_ptr(A), _ptr(W), M, N, K, 3, _ptr(bias), None, None, _ptr(ob), N, _stream()), "gemm qgelu")
    z = ref + bias.float()
    torch.testing.assert_close(ob.float(), z * torch.sigmoid(1.702 * z), rtol=1e-2, atol=1e-2)
    L.check(lib.vdd_gemm(_ptr(A), _ptr(W), M, N, K, 4, _ptr(bias), None, None, _ptr(ob), N, _stream()), "gemm gelu")
    torch.testing.assert_close(ob.float(), torch.nn.functional.gelu(z), rtol=1e-2, atol=1e-2)
    oh = torch.zeros(M, N // 2, device="cuda", dtype=torch.bfloat16)
    L.check(lib.vdd_gemm(_ptr(A), _ptr(W), M, N, K, 5, None, None, None, _ptr(oh), N // 2, _stream()), "gemm swiglu")
    torch.testing.assert_close(oh.float(), torch.nn.functional.silu(ref[:, 0::2]) * ref[:, 1::2], rtol=1e-2, atol=1e-2)
    # row remap (image splice): reversed rows, one skipped
    rm = torch.arange(M - 1, -1, -1, device="cuda", dtype=torch.int32)
    rm[0] = -1
    out = torch.zeros(M, N, device="cuda")
    L.check(lib.vdd_gemm(_ptr(A), _ptr(W), M, N, K, 0, None, _ptr(rm), _ptr(out), None, N, _stream()), "gemm rowmap")
    exp = torch.zeros(M, N, device="cuda")
    exp[rm[1:].long()] = ref[1:]
    torch.testing.assert_close(out, exp, **tol)


# ------------------------------------------------------------------------------------------------ prefill attention
def _attn_reference(q, k, v, causal, scale):
    """fp32 eager attention of one (sequence, head): q [n, D], k / v [n, D] (HF LlamaAttention / CLIPAttention eager path)."""
    s = (q.float() @ k.float().T) * scale
    if causal:
        n = q.shape[0]
        s = s.masked_fill(torch.ones(n, n, dtype=torch.bool, device=q.device).triu(1), float("-inf"))
    return torch.softmax(s, dim=-1) @ v.float()


def _run_prefill_attn(q, q_ld, q_rows, k, v, kv_rows, tok, k_off, hs, q0, lens, out, out_ld, heads, causal, D, scale, use_tc):
    L = _lib()
    lib = L.load()
    dev = q.device
    d_off = torch.tensor(k_off, dtype=torch.int64, device=dev)
    d_hs = torch.tensor(hs, dtype=torch.int64, device=dev)
    d_q0 = torch.tensor(q0, dtype=torch.int32, device=dev)
    d_len = torch.tensor(lens, dtype=torch.int32, device=dev)
    L.check(lib.vdd_debug_prefill_attn(_ptr(q), q_ld, q_rows, _ptr(k), _ptr(v), kv_rows, tok, _ptr(d_off), _ptr(d_hs), _ptr(d_q0), _ptr(d_len),
                                       _ptr(out), out_ld, len(lens), heads, max(lens), int(causal), D, C.c_float(scale), int(use_tc),
                                       _stream()), "prefill_attn")
    torch.cuda.synchronize()


@pytest.mark.parametrize("lens", [[607, 32, 31], [128], [129, 64, 1], [863, 257]])
def test_prefill_attn_tc_llm_layout(lens):
    """tcgen05 + TMA prefill attention on the LLM layout (head_dim 128, causal, keys in a head-major cache with slack rows that
    hold NaN -- the kernel must never let them through) against fp32 eager attention and against the mma.sync kernel."""
    torch.manual_seed(sum(lens))
    dev, heads, D = "cuda", 4, 128
    Hl = heads * D
    Mtot = sum(lens)
    q = (torch.randn(Mtot, Hl, device=dev) * 1.5).bfloat16()
    caps = [n + 40 for n in lens]
    kv_total = sum(c * Hl for c in caps)
    # K slack rows: NaN (scores are masked by select).  V slack rows: finite garbage (the engine zeroes fresh caches; see engine.cu)
    kc = torch.full((kv_total,), float("nan"), device=dev, dtype=torch.bfloat16)
    vc = torch.full((kv_total,), 7.0, device=dev, dtype=torch.bfloat16)
    k_off, hs, q0, ref = [], [], [], torch.zeros(Mtot, Hl, device=dev)
    off = r0 = 0
    for n, cap in zip(lens, caps):
        k_off.append(off); hs.append(cap * D); q0.append(r0)
        for h in range(heads):
            ks = (torch.randn(n, D, device=dev)).bfloat16()
            vs = (torch.randn(n, D, device=dev)).bfloat16()
            base = off + h * cap * D
            kc[base:base + n * D] = ks.flatten()
            vc[base:base + n * D] = vs.flatten()
            ref[r0:r0 + n, h * D:(h + 1) * D] = _attn_reference(q[r0:r0 + n, h * D:(h + 1) * D], ks, vs, True, D ** -0.5)
        off += cap * Hl
        r0 += n
    outs = []
    for use_tc in (1, 0):
        out = torch.full((Mtot, Hl), float("nan"), device=dev, dtype=torch.bfloat16)
        _run_prefill_attn(q, Hl, Mtot, kc, vc, kv_total // 128, 128, k_off, hs, q0, lens, out, Hl, heads, True, D, D ** -0.5, use_tc)
        outs.append(out.float())
    tc, mma = outs
    assert torch.isfinite(tc).all()
    # P is rounded to bf16 before P.V (as the reference rounds to fp16) and the output is bf16: 2e-2 absolute on O(1) values
    assert (tc - ref).abs().max().item() < 2e-2, (tc - ref).abs().max().item()
    assert (tc - mma).abs().max().item() < 2e-2, (tc - mma).abs().max().item()


@pytest.mark.parametrize("B", [1, 3])
def test_prefill_attn_tc_vit_layout(B):
    """The same kernel on the vision tower's layout: head_dim 64, non-causal, q | k | v interleaved in one [rows, 3*VH] buffer."""
    torch.manual_seed(B)
    dev, heads, D, NT = "cuda", 4, 64, 577
    VH = heads * D
    rows = B * NT
    qkv = (torch.randn(rows, 3 * VH, device=dev)).bfloat16()
    ref = torch.zeros(rows, VH, device=dev)
    for b in range(B):
        r = slice(b * NT, (b + 1) * NT)
        for h in range(heads):
            c = slice(h * D, (h + 1) * D)
            ref[r, c] = _attn_reference(qkv[r, :VH][:, c], qkv[r, VH:2 * VH][:, c], qkv[r, 2 * VH:][:, c], False, 0.125)
    k_off = [b * NT * 3 * VH for b in range(B)]
    outs = []
    for use_tc in (1, 0):
        out = torch.full((rows, VH), float("nan"), device=dev, dtype=torch.bfloat16)
        _run_prefill_attn(qkv, 3 * VH, rows, qkv[:, VH:], qkv[:, 2 * VH:], rows, 3 * VH, k_off, [D] * B, [b * NT for b in range(B)],
                          [NT] * B, out, VH, heads, False, D, 0.125, use_tc)
        outs.append(out.float())
    tc, mma = outs
    assert torch.isfinite(tc).all()
    assert (tc - ref).abs().max().item() < 2e-2, (tc - ref).abs().max().item()
    assert (tc - mma).abs().max().item() < 2e-2, (tc - mma).abs().max().item()
