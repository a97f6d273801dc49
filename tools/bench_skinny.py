"""Micro-benchmark of the decode projections for many fused rows (stream-K tcgen05 kernel vs the square-tile GEMM):
weight GB/s per projection shape of LLaVA-1.5-7B / 13B at R rows.  Weights rotate over several buffers so that every
launch streams from HBM (the 126 MB L2 would otherwise keep a 33 MB o_proj resident).
usage: python tools/bench_skinny.py [R ...]"""
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
peak = 6550.0
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass

SHAPES = {"7b": [("qkv", 12288, 4096, 0), ("o", 4096, 4096, 1), ("gate_up", 22016, 4096, 5), ("down", 4096, 11008, 1), ("lm_head", 32000, 4096, 0)],
          "13b": [("qkv", 15360, 5120, 0), ("o", 5120, 5120, 1), ("gate_up", 27648, 5120, 5), ("down", 5120, 13824, 1)]}
rows = [int(x) for x in sys.argv[1:]] or [16, 96, 192]
out = []
for model, shapes in SHAPES.items():
    for name, N, K, mode in shapes:
        nbuf = max(2, int(400e6 // (N * K * 2)) + 1)
        Ws = [(torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16() for _ in range(nbuf)]
        for R in rows:
            x = torch.randn(R, K, device="cuda").bfloat16()
            of = torch.zeros(R, N, device="cuda")
            ob = torch.zeros(R, N // 2, device="cuda", dtype=torch.bfloat16)
            res = {}
            for impl in ("skinny", "square"):
                def run(i):
                    W = Ws[i % nbuf]
                    if impl == "skinny":
                        L.check(lib.vdd_skinny_gemm(P(W), P(x), R, N, K, mode, P(of), P(ob), N // 2 if mode == 5 else N, st()), "skinny")
                    else:
                        L.check(lib.vdd_gemm(P(x), P(W), R, N, K, mode, None, None, P(of), P(ob), N // 2 if mode == 5 else N, st()), "gemm")
                for i in range(3):
                    run(i)
                reps = 30
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for i in range(reps):
                    run(i)
                e1.record()
                torch.cuda.synchronize()
                us = 1e3 * e0.elapsed_time(e1) / reps
                res[impl] = {"us": round(us, 2), "GBps": round(N * K * 2 / us / 1e3, 1), "frac": round(N * K * 2 / us / 1e3 / peak, 3)}
            line = {"model": model, "proj": name, "N": N, "K": K, "R": R, **{k: v for k, v in res.items()}}
            print(json.dumps(line), flush=True)
            out.append(line)
        del Ws
os.makedirs("gpurun_out", exist_ok=True)
json.dump({"peak_GBps": peak, "note": "CUDA events around 30 back-to-back launches (the stand-alone entry adds one 16 KB memset per skinny launch; "
           "the square-tile GEMM has no split-K for ADD through this entry)", "results": out}, open("gpurun_out/bench_skinny.json", "w"), indent=1)
