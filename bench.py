#!/usr/bin/env python
"""Headline benchmark: LLaVA-1.5-7B Visual-Debias-Decoding tokens/s (BASELINE.json config 2: bf16, batch 1, one 336px
image, 32-token prompt, 256 new tokens, use_dd + use_dd_unk => 3 fused streams), random-init weights, synthetic inputs.

  python bench.py --gpus N --steps K --warmup W            # this repo (libvdd)
  python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (oracle port), bounded sample

One "step" = one whole generate() of the workload.  `value` is measured with inputs resident in HBM, `e2e` through
model.generate() with pinned HOST inputs (pixels H2D and tokens D2H inside the timed region).  Prints ONE JSON line.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "LLaVA-1.5-7B VDD decode tokens/sec"
DOMINANT_KERNEL = "vdd::gemv_kernel (decode projections, 3 fused rows; o/gate_up/down/next-qkv chained in one launch per layer)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--model", default="llava-v1.5-7b")
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--prompt-len", type=int, default=32)
    ap.add_argument("--new-tokens", type=int, default=256)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--parallel", default=None, choices=["dp", "tp"],
                    help="N>1 (default tp): tp = ONE request stream sharded tensor-parallel over the N GPUs (heads / FFN columns / vocab; "
                         "all-reduce inside the decode kernels over NVLink peer memory), strong scaling; dp = one independent request "
                         "stream per GPU (how the reference itself uses several GPUs, scripts/pope/run_dataset.sh:14-32), no collective")
    ap.add_argument("--no-configs", action="store_true", help="skip the BASELINE configs[2] (batch 32) block of the JSON line")
    ap.add_argument("--flags", default="both", choices=["both", "unk", "dd"],
                    help="debias streams: both = use_dd + use_dd_unk (3 fused streams, BASELINE configs[1,2,4]); unk = use_dd_unk only "
                         "(2 streams, configs[3]); dd = use_dd only")
    ap.add_argument("--profile", action="store_true", help="under ncu: the device-resident leg only, no extra launches")
    return ap.parse_args()


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                       "-i", str(index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if not self.p:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, reasons, mx = [], set(), None
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f:
            c = [x.strip() for x in line.split(",")]
            if len(c) < 8:
                continue
            try:
                sm.append(float(c[1]))
                mx = float(c[2])
            except ValueError:
                continue
            for n, v in zip(names, c[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if sm:
            sm.sort()
            out = {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return out


def synthetic_inputs(cfg, B, T, seed=0):
    import numpy as np
    import torch
    rng = np.random.default_rng(seed)
    ids = torch.from_numpy(rng.integers(3, cfg.vocab_size, size=(B, T)).astype("int64"))
    ids[:, 1 + T // 8] = -200  # <image> right after the system preamble (llava_sampling.py:69-79)
    img = torch.from_numpy(rng.standard_normal((B, 3, cfg.image_size, cfg.image_size), dtype="float32")).to(torch.bfloat16)
    return ids, img


def algorithmic_bytes(cfg, tp=1):
    """Weights read once per decode step for ALL fused streams."""
    H, I, V, L = cfg.hidden_size, cfg.intermediate_size, cfg.vocab_size, cfg.num_hidden_layers
    per_layer = 4 * H * H + 3 * H * I
    proj = 2 * (L * per_layer + V * H) // tp
    kv_tok = L * 2 * H * 2 // tp
    return proj, kv_tok


def run_reference(a):
    """Reference arm: the reference's CPU path (oracle port, FULL depth) on the host cores.  One step = one bounded sample of
    the workload (image pass + the three prefills + 3 decode steps, all measured); `ms_per_step` is the measured wall time of
    a sample, `value` the tokens/s of the whole 256-token workload projected from the measured phases."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from llava_align_b200.config import by_name
    from oracle import cpu_baseline as cb
    cfg = by_name(a.model)
    vals, walls, last = [], [], None
    sample = cb.ReferenceSample(cfg, B=a.batch, T=a.prompt_len, n_new=4)  # built once; each step is one bounded generate
    t_start = time.perf_counter()
    for i in range(a.warmup + a.steps):
        r = sample.run()
        if i >= a.warmup:
            vals.append(cb.tokens_per_s(r, a.batch, a.new_tokens))
            walls.append(r["wall_s"])
            last = r
        if time.perf_counter() - t_start > 200 and vals:
            break
    v = sum(vals) / len(vals)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "tokens/s", "n_gpus": a.gpus, "steps": len(vals),
            "warmup": a.warmup, "ms_per_step": 1000.0 * sum(walls) / len(walls), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "value_kind": "projected: B*new_tokens / (vision_s + prefill_s + (new_tokens-1)*decode_s_per_token), every phase measured at "
                          "full model depth inside the step; ms_per_step is the measured wall time of one bounded sample",
            "projected_ms_per_full_generate": 1000.0 * a.batch * a.new_tokens / v,
            "config": workload_config(a, cfg, 1),
            "cpu_baseline": {"value": v, "unit": "tokens/s", "cores": last["cores"], "kind": "port", "sample": last["sample"],
                             "phases": {k: last[k] for k in ("vision_s", "prefill_s", "decode_s_per_token")},
                             "build_s": last["measured"]["build_s"]},
            "e2e": {"value": v, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    _emit(line)


def stream_flags(a):
    """(use_dd_unk, use_dd, number of fused streams, text)"""
    return {"both": (True, True, 3, "use_dd+use_dd_unk (3 fused streams)"), "unk": (True, False, 2, "use_dd_unk (2 fused streams)"),
            "dd": (False, True, 2, "use_dd (2 fused streams)")}[a.flags]


def workload_config(a, cfg, n, mode="dp"):
    is13 = "13b" in cfg.name
    if not is13 and a.batch == 1 and a.prompt_len == 32 and a.new_tokens == 256 and a.flags == "both":
        which = "BASELINE.json configs[1]"
    elif not is13 and a.batch == 32 and a.flags == "both":
        which = "BASELINE.json configs[2] shape (POPE-like batch of 32)"
    elif is13 and a.batch == 8 and a.flags == "unk":
        which = "BASELINE.json configs[3] shape (13B, batch 8, use_dd_unk)"
    elif is13 and a.batch == 64 and a.flags == "both":
        which = "BASELINE.json configs[4] shape (13B, batch 64, both dd streams)"
    else:
        which = "custom shape (not a BASELINE.json config)"
    return {"workload": f"{which}: {cfg.name} bf16 random-init, batch {a.batch}, 1 synthetic "
                        f"{cfg.image_size}px image, {a.prompt_len}-token prompt, {a.new_tokens} new tokens, {stream_flags(a)[3]}, "
                        f"cd_alpha=1 cd_beta=0.1, temperature-1 sampling, EOS disabled",
            "global_batch": a.batch, "new_tokens": a.new_tokens, "prompt_len": a.prompt_len, "streams": stream_flags(a)[2],
            "parallelism": "single GPU" if n == 1 else (
                f"tp{n}: ONE request stream sharded over {n} GPUs (heads / FFN columns / vocab split; decode: all-reduce and logit "
                f"all-gather inside the projection kernels over NVLink peer memory, no NCCL call; prefill: NCCL all-reduce)" if mode == "tp"
                else f"dp{n}: {n} independent replicas, one request stream per GPU, no data-path collective"),
            "l2_policy": "inputs larger than L2: every step streams the 13.2 GB of weights (L2 = 126 MB); no flush needed"}


def _emit(line: dict):
    """The ONE JSON line goes to the real stdout; everything else any library prints (e.g. NCCL's version banner) was
    diverted to stderr by _divert_stdout()."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def _divert_stdout():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def main():
    a = parse()
    _divert_stdout()
    if a.profile:
        a.no_e2e = a.no_cpu_baseline = True
    if a.impl == "reference":
        return run_reference(a)
    import torch
    import torch.distributed as dist
    from llava_align_b200 import _lib
    from llava_align_b200 import engine as E
    from llava_align_b200.builder import load_pretrained_model
    from llava_align_b200.config import by_name
    from llava_align_b200.model import evolve_vcd_sampling

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = by_name(a.model)
    evolve_vcd_sampling()
    if a.parallel is None:
        a.parallel = "tp" if world > 1 else "dp"
    tp = world if a.parallel == "tp" else 1
    tok, model, _, _ = load_pretrained_model(f"random:{a.model}:0", None, a.model, device=str(dev), tp_size=tp,
                                             tp_rank=rank if tp > 1 else 0)
    eng = model.engine
    if tp > 1:
        eng.tp_setup(dist)
    V = cfg.vocab_size
    L = cfg.num_hidden_layers
    lib = _lib.load()
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
    else:
        peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
    proj_bytes, kv_tok = algorithmic_bytes(cfg, tp)
    replicas = world if tp == 1 else 1  # dp: every rank generated its own B x n_new tokens in the (max-over-ranks) time

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, warm, steps):
        for _ in range(warm):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        eng.launch_count(reset=True)
        with torch.cuda.stream(eng.stream):
            e0.record(eng.stream)
            for _ in range(steps):
                fn()
            e1.record(eng.stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, eng.launch_count()

    def note(msg):
        print(f"[bench rank {rank}] {msg}", file=sys.stderr, flush=True)

    def measure(B, T, n_new, warm, steps, with_e2e):
        """One workload: device-resident tokens/s, end-to-end tokens/s through model.generate with pinned host inputs, the
        decode step against its own algorithmic bytes, and the projection launches of one step timed alone."""
        ids, img = synthetic_inputs(cfg, B, T)
        img_pinned, img_dev, ids_pinned = img.pin_memory(), img.to(dev), ids.pin_memory()
        params = E.make_sample_params(cd_alpha=1.0, cd_beta=0.1, use_contrast=True, eos_token_id=None)
        g = torch.Generator(device=dev).manual_seed(1234)
        q = torch.empty(n_new, B, V, device=dev).exponential_(1, generator=g)  # identical on every TP rank
        use_unk, use_none, S, _ = stream_flags(a)
        dev_fn = lambda: eng.generate(ids, img_dev, use_unk, use_none, params, n_new, q, None)
        ms, launches = timed(dev_fn, warm, steps)
        tm = eng.last_timings()
        out = {"value": replicas * B * n_new * steps / (ms / 1e3), "ms_per_step": ms / steps, "launches": int(launches) * replicas,
               "e2e": None}
        if with_e2e:
            gen_kw = dict(images=img_pinned, images_cd=None, cd_alpha=1.0, cd_beta=0.1, use_dd=use_none, use_dd_unk=use_unk, do_sample=True,
                          temperature=1.0, top_p=None, top_k=None, max_new_tokens=n_new, use_cache=True, eos_token_id=None)
            e2e_fn = lambda: model.generate(ids_pinned, **gen_kw)  # TP: generate() broadcasts rank 0's variates itself
            ms2, _ = timed(e2e_fn, 1, steps)
            out["e2e"] = {"value": replicas * B * n_new * steps / (ms2 / 1e3), "unit": "tokens/s",
                          "h2d_bytes_per_step": int(replicas * (img.numel() * 2 + ids.numel() * 8)),
                          "d2h_bytes_per_step": int(replicas * B * n_new * 8), "ms_per_step": ms2 / steps}
        # projection launches of ONE decode step, replayed back to back (CUDA events on the launching stream)
        reps = 1 if a.profile else 20
        with torch.cuda.stream(eng.stream):
            for _ in range(0 if a.profile else 3):
                _lib.check(lib.vdd_bench_projection_chain(eng.h, eng.stream.cuda_stream), "chain")
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            eng.launch_count(reset=True)
            e0.record(eng.stream)
            for _ in range(reps):
                _lib.check(lib.vdd_bench_projection_chain(eng.h, eng.stream.cuda_stream), "chain")
            e1.record(eng.stream)
        torch.cuda.synchronize()
        out["chain_launches"] = max(1, eng.launch_count() // reps)
        out["chain_ms"] = max(e0.elapsed_time(e1) / reps, 1e-6)
        out["chain_reps"] = reps
        # whole decode step against its own algorithmic bytes (weights once + KV of all rows at mid-run + logits)
        ctx_mid = (T + cfg.num_patches - 1) + (T if use_unk else 0) + ((T - 1) if use_none else 0) + S * (n_new // 2)
        step_bytes = proj_bytes + B * ctx_mid * kv_tok + S * B * V * 4 // tp
        step_ms = tm["decode_ms"] / max(1, n_new - 1)
        out["decode_step"] = {"ms_per_token_step": step_ms, "tokens_per_s_decode_only": B * 1e3 / step_ms if step_ms > 0 else None,
                              "algorithmic_GB_per_step": step_bytes / 1e9,
                              "achieved_GBps": step_bytes / (step_ms / 1e3) / 1e9 if step_ms > 0 else None,
                              "frac_of_peak": (step_bytes / (step_ms / 1e3) / 1e9 / peak) if step_ms > 0 else None,
                              "vision_ms": tm["vision_ms"], "prefill_ms": tm["prefill_ms"]}
        del q, img_dev
        return out

    note(f"model loaded (world {world}, parallel {a.parallel}, tp {tp})")
    sampler = ClockSampler(local) if rank == 0 else None
    B, T, n_new = a.batch, a.prompt_len, a.new_tokens
    m = measure(B, T, n_new, a.warmup, a.steps, not a.no_e2e)
    clocks = sampler.stop() if sampler else None
    note(f"headline workload done: {m['ms_per_step']:.1f} ms per generate")
    value, ms, launches, e2e, decode = m["value"], m["ms_per_step"] * a.steps, m["launches"], m["e2e"], m["decode_step"]

    # ---- roofline of the dominant kernel (the decode projections), timed live with CUDA events
    n_proj = 4 * L + 1
    S_streams = stream_flags(a)[2]
    fused_rows = S_streams * B <= 8  # the GEMV chain is the R <= 8 path; larger batches decode through gemm_tc_kernel
    n_gemv, chain_ms, reps = m["chain_launches"], m["chain_ms"], m["chain_reps"]
    achieved = proj_bytes / (chain_ms / 1e3) / 1e9
    traffic = None
    traffic_path = os.path.join(ROOT, "profiles", "gemv_traffic.json")
    if os.path.exists(traffic_path) and fused_rows and tp == 1:
        traffic = json.load(open(traffic_path)).get("dram_bytes_per_launch")
    roofline = {"bound": "hbm", "kernel": DOMINANT_KERNEL, "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "bytes_per_launch": proj_bytes / n_gemv, "us_per_launch": 1e3 * chain_ms / n_gemv, "launches_per_step": n_gemv,
                "projections_per_step": n_proj,
                "how": f"CUDA events around {reps} back-to-back replays of one decode step's {n_gemv} projection launches "
                       f"({n_proj} projections, {proj_bytes / 1e9:.2f} GB of weights per GPU)"}
    if not fused_rows:
        roofline.update({"kernel": f"vdd::gemm_tc_kernel (tcgen05; decode projections of {S_streams * B} fused rows: qkv, o, gate_up, down per layer + "
                                   f"lm_head, with the split-K reduce + RMSNorm passes between them)", "traffic": None,
                         "how": f"CUDA events around {reps} back-to-back replays of one decode step's projection launches "
                                f"({n_gemv} launches incl. {2 * L + 1} norm passes, {proj_bytes / 1e9:.2f} GB of weights per GPU)"})

    # ---- BASELINE.json configs[2]: the throughput shape (batch 32 POPE-like prompts -> 96 fused rows), same model, same line
    configs = None
    if world == 1 and not a.no_configs and not a.profile and B == 1:
        c3 = measure(32, 64, 64, 1, 3, not a.no_e2e)
        configs = {"configs[2]": {
            "workload": f"BASELINE.json configs[2]: {cfg.name} bf16, batch 32 POPE-shape prompts (64 tokens incl. <image>), 64 new tokens, "
                        f"use_dd+use_dd_unk (96 fused rows), cd_alpha=1 cd_beta=0.1, temperature-1 sampling",
            "value": c3["value"], "unit": "tokens/s", "ms_per_step": c3["ms_per_step"], "steps": 3, "warmup": 1, "e2e": c3["e2e"],
            "decode_step": c3["decode_step"], "gpu_launches": c3["launches"],
            "roofline": {"bound": "hbm", "kernel": "vdd::gemm_tc_kernel<deep> (tcgen05 decode projections of 96 fused rows)",
                         "achieved": proj_bytes / (c3["chain_ms"] / 1e3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": proj_bytes / (c3["chain_ms"] / 1e3) / 1e9 / peak, "launches_per_step": c3["chain_launches"]}}}
        note(f"configs[2] done: {c3['ms_per_step']:.1f} ms per generate")

    # ---- the second half of BASELINE.json's metric: logit-combine GB/s (K8 alone, 3 streams + q, scores not requested).
    # 200 launches captured in one CUDA graph and replayed: per-launch time of the KERNEL (a Python loop would time ctypes).
    combine = None
    if rank == 0 and not a.profile:
        combine = []
        pc = E.make_sample_params(cd_alpha=1.0, cd_beta=0.1, use_contrast=True, eos_token_id=None)
        for Bc in (1, 32, 64, 1024):
            gl = torch.Generator(device=dev).manual_seed(Bc)
            l = torch.randn(Bc, V, device=dev, generator=gl) * 2.5
            u = l + torch.randn(Bc, V, device=dev, generator=gl)
            d = l + torch.randn(Bc, V, device=dev, generator=gl)
            qq = torch.empty(Bc, V, device=dev).exponential_(1, generator=gl)
            toks = torch.empty(Bc, dtype=torch.int64, device=dev)
            torch.cuda.synchronize()
            ms_c = C.c_float(0)
            reps_c = 200 if Bc <= 64 else 20
            _lib.check(lib.vdd_bench_combine(l.data_ptr(), u.data_ptr(), d.data_ptr(), Bc, V, V, C.byref(pc), qq.data_ptr(),
                                             toks.data_ptr(), reps_c, C.byref(ms_c), eng.stream.cuda_stream), "vdd_bench_combine")
            us = 1e3 * ms_c.value
            nbytes = Bc * V * 4 * 4 + Bc * 8
            # the same launch as generate() issues it: no q tensor, the variates drawn in-kernel (Philox, torch's stream)
            _lib.check(lib.vdd_bench_combine(l.data_ptr(), u.data_ptr(), d.data_ptr(), Bc, V, V, C.byref(pc), None,
                                             toks.data_ptr(), reps_c, C.byref(ms_c), eng.stream.cuda_stream), "vdd_bench_combine")
            us_p = 1e3 * ms_c.value
            nbytes_p = Bc * V * 4 * 3 + Bc * 8
            combine.append({"B": Bc, "us_per_launch": us, "GBps": nbytes / us / 1e3, "bytes": nbytes, "frac_of_hbm_peak": nbytes / us / 1e3 / peak,
                            "resident": "L2 (same inputs every replay)" if nbytes < 100e6 else "HBM (inputs larger than L2)",
                            "how": f"{reps_c} launches in one CUDA graph, best of 3 replays",
                            "in_kernel_variates": {"us_per_launch": us_p, "bytes": nbytes_p, "GBps": nbytes_p / us_p / 1e3,
                                                   "frac_of_hbm_peak": nbytes_p / us_p / 1e3 / peak}})
        del l, u, d, qq

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        from oracle import cpu_baseline as cb
        r = cb.ReferenceSample(cfg, B=B, T=T, n_new=4).run()
        cpu = {"value": cb.tokens_per_s(r, B, n_new), "unit": "tokens/s", "cores": r["cores"], "kind": "port", "sample": r["sample"],
               "value_kind": "projected from the measured full-depth phases of one bounded sample", "sample_wall_s": r["wall_s"],
               "phases": {k: r[k] for k in ("vision_s", "prefill_s", "decode_s_per_token")}}
    note("all legs done")
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "tokens/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "strong" if tp > 1 else "weak",
                "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": workload_config(a, cfg, world, a.parallel),
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "decode_step": decode,
                "configs": configs, "logit_combine": combine, "cpu_baseline": cpu}
        _emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
