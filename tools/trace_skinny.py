"""Where the time of one stream-K decode launch goes: per-CTA %globaltimer stamps (vdd_skinny_gemm_trace) of a single
projection and of a whole layer chain.  Prints min/median/max over CTAs of every stamp relative to the earliest kernel
entry (ns) and writes gpurun_out/trace_skinny.json"""
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from llava_align_b200 import _lib as L  # noqa: E402

lib = L.load()
P = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
NAMES = ["entry", "w_requested", "prev_done", "first_stage", "last_mma", "p0_drained", "p0_fixed_arrive", "p0_barrier_seen",
         "p0_first", "p0_last_mma", "p1_first", "p1_last_mma", "p2_first", "p2_last_mma", "p3_first", "p3_last_mma"]


def summarize(buf, tag):
    t = buf.cpu().view(148, 16)
    t = t[t[:, 0] > 0]
    t0 = int(t[:, 0].min())
    line = dict(tag)
    line["ctas"] = int(t.shape[0])
    for j, nm in enumerate(NAMES):
        col = t[:, j]
        col = col[col > 0] - t0
        if len(col):
            line[nm] = {"min": int(col.min()), "med": int(col.median()), "max": int(col.max())}
    print(json.dumps(line), flush=True)
    return line


out = []
mk = lambda n, k: (torch.randn(n, k, device="cuda") / k ** 0.5).bfloat16()
for name, N, K, mode, R in [("o", 4096, 4096, 1, 96), ("gate_up", 22016, 4096, 5, 96), ("qkv", 12288, 4096, 0, 96)]:
    Ws = [mk(N, K) for _ in range(3)]
    x = torch.randn(R, K, device="cuda").bfloat16()
    of = torch.zeros(R, N, device="cuda")
    ob = torch.zeros(R, N // 2, device="cuda", dtype=torch.bfloat16)
    ld = N // 2 if mode == 5 else N
    for i in range(3):
        L.check(lib.vdd_skinny_gemm(P(Ws[i]), P(x), R, N, K, mode, P(of), P(ob), ld, st()), "warm")
    torch.cuda.synchronize()
    buf = torch.zeros(148 * 16, dtype=torch.int64, device="cuda")
    lib.vdd_skinny_gemm_trace(P(buf))
    L.check(lib.vdd_skinny_gemm(P(Ws[0]), P(x), R, N, K, mode, P(of), P(ob), ld, st()), "traced")
    torch.cuda.synchronize()
    lib.vdd_skinny_gemm_trace(None)
    out.append(summarize(buf, {"what": "single projection", "proj": name, "N": N, "K": K, "R": R}))
    del Ws
H, I = 4096, 11008
for R in (96, 16):
    layers = [dict(wo=mk(H, H), wgu=mk(2 * I, H), wd=mk(H, I), wq=mk(3 * H, H)) for _ in range(2)]
    g = torch.ones(H, device="cuda").bfloat16()
    attn = torch.randn(R, H, device="cuda").bfloat16()
    resid = torch.randn(R, H, device="cuda")
    xg = torch.empty(R, H, device="cuda", dtype=torch.bfloat16)
    hb = torch.empty(R, I, device="cuda", dtype=torch.bfloat16)
    qkv = torch.empty(R, 3 * H, device="cuda")
    run = lambda w: L.check(lib.vdd_skinny_layer_chain(P(attn), P(resid), P(w["wo"]), P(g), P(w["wgu"]), P(w["wd"]), P(g), P(w["wq"]), R, H, H, I,
                                                       3 * H, 1e-5, P(xg), P(hb), P(qkv), st()), "chain")
    for i in range(3):
        run(layers[i % 2])
    torch.cuda.synchronize()
    buf = torch.zeros(148 * 16, dtype=torch.int64, device="cuda")
    lib.vdd_skinny_gemm_trace(P(buf))
    run(layers[1])
    torch.cuda.synchronize()
    lib.vdd_skinny_gemm_trace(None)
    out.append(summarize(buf, {"what": "7B layer chain (o, gate_up, down, qkv): 404.8 MB", "R": R}))
    del layers
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/trace_skinny.json", "w"), indent=1)
