"""Debug tool: per-CTA time stamps of one fused projection chain (middle layer) during a real decode step.
VDD_TRACE=1 python tools/trace_chain.py"""
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
n = 148 * 4 * 8
buf = (C.c_ulonglong * n)()
_lib.check(lib.vdd_debug_trace(m.engine.h, buf, n), "trace")
raw = np.array(buf[:], dtype=np.int64).reshape(148, 4, 8)
smid = raw[:, 0, 6].copy()
t = raw.astype(np.float64)
t0 = t[:, 0, 0].min()
t = (t - t0) / 1000.0  # us
names = ["o_proj", "gate_up", "down", "qkv_next"]
ev = ["kernel_start", "phase_start(after barrier)", "x staged", "loop done (thread0)", "all warps done", "epilogue done"]
out = {}
for p in range(4):
    print(f"--- phase {p} {names[p]}")
    for e in range(6):
        if p > 0 and e == 0:
            continue
        col = t[:, p, e]
        print(f"  {ev[e]:28s} min {col.min():8.2f}  mean {col.mean():8.2f}  max {col.max():8.2f} us")
        out[f"{names[p]}.{ev[e]}"] = [float(col.min()), float(col.mean()), float(col.max())]
dur = t[:, :, 3] - t[:, :, 2]  # stream-loop duration per CTA and phase
print("loop duration per phase: mean / std / min / max")
for p in range(4):
    print(f"  {names[p]:9s} {dur[:, p].mean():6.2f} {dur[:, p].std():5.2f} {dur[:, p].min():6.2f} {dur[:, p].max():6.2f}")
print("correlation of per-CTA loop durations between phases (systematic per-SM speed?)")
print(np.round(np.corrcoef(dur.T), 2))
order = np.argsort(dur[:, 1])
print("slowest CTAs in gate_up (cta, smid, dur gate_up, dur down, dur qkv):")
for i in order[-8:]:
    print("  ", int(i), int(smid[i]), round(dur[i, 1], 2), round(dur[i, 2], 2), round(dur[i, 3], 2))
print("fastest:")
for i in order[:8]:
    print("  ", int(i), int(smid[i]), round(dur[i, 1], 2), round(dur[i, 2], 2), round(dur[i, 3], 2))
ep = t[:, :, 5] - t[:, :, 4]
mid = t[:, :, 7] - t[:, :, 4]
print("epilogue total / up-to-partial-sum (thread 0), mean per phase:", np.round(ep.mean(0), 2), np.round(mid.mean(0), 2))
out["loop_dur_mean"] = dur.mean(0).tolist(); out["loop_dur_std"] = dur.std(0).tolist()
out["corr"] = np.corrcoef(dur.T).tolist()
json.dump(out, open("gpurun_out/trace_chain.json", "w"), indent=1)
