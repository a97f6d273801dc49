"""Tuning tool: time one decode step's projection chain for several warp counts (VDD_GV_WARPS)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from llava_align_b200 import _lib, engine as E
from llava_align_b200.builder import load_pretrained_model
from llava_align_b200.model import evolve_vcd_sampling
from bench import synthetic_inputs
evolve_vcd_sampling()
_, m, _, _ = load_pretrained_model("random:llava-v1.5-7b:0", None, "llava-v1.5-7b")
eng = m.engine
ids, img = synthetic_inputs(m.vdd_config, 1, 32)
params = E.make_sample_params(cd_alpha=1.0, cd_beta=0.1, use_contrast=True, eos_token_id=None)
eng.generate(ids, img.cuda(), True, True, params, 8, None, None)
lib = _lib.load()
for w in [8, 6, 4, 3, 2, 1]:
    os.environ["VDD_GV_KSMUL"] = str(w)
    with torch.cuda.stream(eng.stream):
        for _ in range(3):
            _lib.check(lib.vdd_bench_projection_chain(eng.h, eng.stream.cuda_stream), "chain")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(eng.stream)
        for _ in range(20):
            _lib.check(lib.vdd_bench_projection_chain(eng.h, eng.stream.cuda_stream), "chain")
        e1.record(eng.stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"ksmul={w}: chain {ms:.3f} ms -> {13.2147 / ms:.1f} GB/ms = {13214.7 / ms / 6548.2:.3f} of peak", flush=True)
