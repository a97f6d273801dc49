"""Debug tool: per-CTA %globaltimer stamps of the first 64 phases of the whole-step kernel (decode_step.cu) inside a real
decode step.  VDD_TRACE=1 python tools/trace_step.py [out.json]
Stamps per (cta, phase): 1 phase start (after the previous grid barrier), 2 x staged, 3 stream loop done (thread 0),
4 all warps done, 5 epilogue done.  Attention phases only carry stamp 1."""
import ctypes as C, json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["VDD_TRACE"] = "1"
from llava_align_b200 import _lib, engine as E
from llava_align_b200.builder import load_pretrained_model
from llava_align_b200.model import evolve_vcd_sampling
from bench import synthetic_inputs
evolve_vcd_sampling()
_, m, _, _ = load_pretrained_model("random:llava-v1.5-7b:0", None, "llava-v1.5-7b")
cfg = m.vdd_config
ids, img = synthetic_inputs(cfg, 1, 32)
params = E.make_sample_params(cd_alpha=1.0, cd_beta=0.1, use_contrast=True, eos_token_id=None)
for _ in range(2):
    m.engine.generate(ids, img.cuda(), True, True, params, 64, None, None)
lib = _lib.load()
G, P = 148, 64
n = G * P * 8
buf = (C.c_ulonglong * n)()
_lib.check(lib.vdd_debug_trace(m.engine.h, buf, n), "trace")
t = np.array(buf[:], dtype=np.float64).reshape(G, P, 8)
t0 = t[:, 0, 1].min()
t = (t - t0) / 1000.0  # us
names = ["qkv0"] + [f"L{l}.{n}" for l in range(13) for n in ("attn", "o", "gate_up", "down", "qkv")]
out = {}
print(f"{'phase':12s} {'start(min/mean/max)':>26s} {'x staged':>9s} {'loop done':>10s} {'epi done':>9s} | {'prologue':>8s} {'loop':>7s} {'epi':>6s} {'gap to next':>11s}")
for p in range(1, 32):
    st = t[:, p, 1]
    row = {"start": [st.min(), st.mean(), st.max()]}
    if "attn" in names[p]:
        nxt = t[:, p + 1, 1]
        print(f"{names[p]:12s} {st.min():8.2f} {st.mean():8.2f} {st.max():8.2f}   (attention phase: {nxt.mean() - st.mean():6.2f} us to the next phase start)")
        row["dur"] = float(nxt.mean() - st.mean())
    else:
        xs, ld, ep, nxt = t[:, p, 2], t[:, p, 4], t[:, p, 5], t[:, p + 1, 1]
        print(f"{names[p]:12s} {st.min():8.2f} {st.mean():8.2f} {st.max():8.2f} {xs.mean():9.2f} {ld.mean():10.2f} {ep.mean():9.2f} | "
              f"{(xs - st).mean():8.2f} {(ld - xs).mean():7.2f} {(ep - ld).mean():6.2f} {(nxt - ep).mean():11.2f}")
        row.update(prologue=float((xs - st).mean()), loop=float((ld - xs).mean()), epilogue=float((ep - ld).mean()),
                   barrier=float((nxt - ep).mean()), loop_spread=float(ld.max() - ld.min()))
    out[names[p]] = row
layer = t[:, 26, 1].mean() - t[:, 21, 1].mean()
print(f"one layer (L4.attn start -> L5.attn start): {layer:.2f} us")
out["layer_us"] = float(layer)
path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/trace_step.json"
json.dump(out, open(path, "w"), indent=1)
