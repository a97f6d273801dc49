"""Where the time of one stream-K decode GEMM launch goes: per-CTA %globaltimer stamps (vdd_skinny_gemm_trace).
usage: python tools/trace_skinny.py  -> prints, per shape, min/median/max over CTAs of every stamp relative to the
earliest kernel entry (ns), and writes gpurun_out/trace_skinny.json"""
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
NAMES = ["entry", "w_requested", "prev_done", "first_stage", "last_mma", "drained", "shared_ready", "exit"]
cases = [("o", 4096, 4096, 1, 96), ("gate_up", 22016, 4096, 5, 96), ("qkv", 12288, 4096, 0, 96), ("qkv", 12288, 4096, 0, 16),
         ("down", 4096, 11008, 1, 96), ("lm_head", 32000, 4096, 0, 192)]
if len(sys.argv) > 1:
    cases = [c for c in cases if c[0] in sys.argv[1:]]
out = []
for name, N, K, mode, R in cases:
    Ws = [(torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16() for _ in range(3)]
    x = torch.randn(R, K, device="cuda").bfloat16()
    of = torch.zeros(R, N, device="cuda")
    ob = torch.zeros(R, N // 2, device="cuda", dtype=torch.bfloat16)
    ld = N // 2 if mode == 5 else N
    for i in range(3):
        L.check(lib.vdd_skinny_gemm(P(Ws[i]), P(x), R, N, K, mode, P(of), P(ob), ld, st()), "warm")
    torch.cuda.synchronize()
    buf = torch.zeros(148 * 8, dtype=torch.int64, device="cuda")
    lib.vdd_skinny_gemm_trace(P(buf))
    L.check(lib.vdd_skinny_gemm(P(Ws[0]), P(x), R, N, K, mode, P(of), P(ob), ld, st()), "traced")
    torch.cuda.synchronize()
    lib.vdd_skinny_gemm_trace(None)
    t = buf.cpu().view(148, 8)
    live = t[:, 0] > 0
    t = t[live]
    t0 = int(t[:, 0].min())
    rel = (t - t0).float()
    rel[t == 0] = float("nan")
    line = {"proj": name, "N": N, "K": K, "R": R, "ctas": int(live.sum())}
    for j, nm in enumerate(NAMES):
        col = rel[:, j]
        col = col[~torch.isnan(col)]
        if len(col):
            line[nm] = {"min": int(col.min()), "med": int(col.median()), "max": int(col.max())}
    print(json.dumps(line), flush=True)
    out.append(line)
    del Ws
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/trace_skinny.json", "w"), indent=1)
