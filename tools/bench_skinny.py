"""Micro-benchmark of the many-row decode projections: one decoder layer's chain (o_proj, gate_up, down_proj, next qkv) as
ONE stream-K tcgen05 launch (vdd_skinny_layer_chain) vs the same four projections as separate square-tile GEMM launches
(vdd_gemm; RMSNorm launches not included on that side).  Weights rotate over several layer copies so that every launch
streams from HBM.  usage: python tools/bench_skinny.py [R ...]"""
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
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
peak = 6550.0
try:
    peak = json.load(open(os.path.join(root, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass

MODELS = {"7b": (4096, 11008), "13b": (5120, 13824)}
rows = [int(x) for x in sys.argv[1:]] or [16, 96, 128]
out = []
for model, (H, I) in MODELS.items():
    nl = 3  # layer copies: 3 x 405 MB (7B) >> L2
    mk = lambda n, k: (torch.randn(n, k, device="cuda") / k ** 0.5).bfloat16()
    layers = [dict(wo=mk(H, H), wgu=mk(2 * I, H), wd=mk(H, I), wq=mk(3 * H, H), g2=torch.ones(H, device="cuda").bfloat16(),
                   g3=torch.ones(H, device="cuda").bfloat16()) for _ in range(nl)]
    layer_bytes = 2 * (H * H + 2 * I * H + H * I + 3 * H * H)
    for R in rows:
        attn = torch.randn(R, H, device="cuda").bfloat16()
        resid = torch.randn(R, H, device="cuda")
        xg = torch.empty(R, H, device="cuda", dtype=torch.bfloat16)
        hb = torch.empty(R, I, device="cuda", dtype=torch.bfloat16)
        qkv = torch.empty(R, 3 * H, device="cuda")

        def chain(i):
            w = layers[i % nl]
            L.check(lib.vdd_skinny_layer_chain(P(attn), P(resid), P(w["wo"]), P(w["g2"]), P(w["wgu"]), P(w["wd"]), P(w["g3"]), P(w["wq"]),
                                               R, H, H, I, 3 * H, 1e-5, P(xg), P(hb), P(qkv), st()), "chain")

        def square(i):
            w = layers[i % nl]
            L.check(lib.vdd_gemm(P(attn), P(w["wo"]), R, H, H, 1, None, None, P(resid), None, H, st()), "o")
            L.check(lib.vdd_gemm(P(xg), P(w["wgu"]), R, 2 * I, H, 5, None, None, None, P(hb), I, st()), "gu")
            L.check(lib.vdd_gemm(P(hb), P(w["wd"]), R, H, I, 1, None, None, P(resid), None, H, st()), "down")
            L.check(lib.vdd_gemm(P(xg), P(w["wq"]), R, 3 * H, H, 0, None, None, P(qkv), None, 3 * H, st()), "qkv")

        res = {}
        for name, fn in (("chain", chain), ("square", square)):
            for i in range(3):
                fn(i)
                resid.normal_()
            reps = 24
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(reps):
                fn(i)
            e1.record()
            torch.cuda.synchronize()
            us = 1e3 * e0.elapsed_time(e1) / reps
            res[name] = {"us_per_layer": round(us, 2), "GBps": round(layer_bytes / us / 1e3, 1), "frac": round(layer_bytes / us / 1e3 / peak, 3)}
            resid.normal_()
        line = {"model": model, "R": R, "layer_MB": round(layer_bytes / 1e6, 1), **res}
        print(json.dumps(line), flush=True)
        out.append(line)
    del layers
os.makedirs(os.path.join(root, "gpurun_out"), exist_ok=True)
json.dump({"peak_GBps": peak, "note": "CUDA events around 24 back-to-back layers; chain = one launch (+ one 4 KB memset) per layer, "
           "square = four launches per layer (no split-K and no RMSNorm launches through this entry)", "results": out},
          open(os.path.join(root, "gpurun_out", "bench_skinny.json"), "w"), indent=1)
