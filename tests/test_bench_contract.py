"""CPU checks of bench.py's contract: the algorithmic-byte model behind `roofline` / `decode_step` against SURVEY.md 8(d), and
the key set of the committed bench lines (profiles/) against what the driver reads."""
import importlib.util
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_algorithmic_bytes_match_the_survey():
    """7B  W = 13,214,687,232 B, KVtok = 524,288 B;  13B  W = 25,704,048,640 B, KVtok = 819,200 B
    (bench.py leaves out the RMSNorm vectors: 2H per layer + H, < 0.01 %)."""
    from llava_align_b200.config import LLAVA_15_7B, LLAVA_15_13B
    b = _bench()
    w7, kv7 = b.algorithmic_bytes(LLAVA_15_7B)
    w13, kv13 = b.algorithmic_bytes(LLAVA_15_13B)
    assert abs(w7 - 13_214_687_232) / 13_214_687_232 < 1e-4 and kv7 == 524_288
    assert abs(w13 - 25_704_048_640) / 25_704_048_640 < 1e-4 and kv13 == 819_200
    # tensor parallel: every rank streams 1/t of both
    w7t, kv7t = b.algorithmic_bytes(LLAVA_15_7B, tp=8)
    assert w7t == w7 // 8 and kv7t == kv7 // 8
    # the worked ceiling of SURVEY 8(d) for config 2 mid-run: 13.77 GB per step
    ctx_mid = (32 + 576 - 1) + 32 + 31 + 3 * 128
    assert abs((w7 + ctx_mid * kv7 + 3 * 32000 * 4) / 1e9 - 13.77) < 0.01


@pytest.mark.parametrize("name", ["r1_bench_n1_latest.json", "r1_bench_config3_b32.json", "r1_bench_dp2.json", "r1_bench_tp2_fused.json",
                                  "r2_bench_n1.json", "r2_bench_tp2.json", "r2_bench_tp4.json", "r2_bench_tp8.json",
                                  "r2_bench_config3_13b_b8_tp4.json"])
def test_committed_bench_lines_carry_the_contract_keys(name):
    d = json.load(open(os.path.join(ROOT, "profiles", name)))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "clocks", "gpu_launches", "roofline"):
        assert k in d, k
    assert d["unit"] == "tokens/s" and d["higher_is_better"] is True and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"]
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert d["gpu_launches"] > 0 and d["value"] > 0
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    bad = {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(d["clocks"]["reasons"])
    assert not bad, bad
    if d.get("e2e"):
        assert set(d["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} and d["e2e"]["h2d_bytes_per_step"] > 0


def test_reference_arm_line():
    d = json.load(open(os.path.join(ROOT, "profiles", "r1_bench_reference_arm.json")))
    assert d["impl"] == "reference" and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]


def test_round2_line_carries_configs_block_and_tp_default():
    d = json.load(open(os.path.join(ROOT, "profiles", "r2_bench_n1.json")))
    c = d["configs"]["configs[2]"]
    for k in ("workload", "value", "unit", "e2e", "decode_step", "roofline", "gpu_launches"):
        assert k in c, k
    assert "batch 32" in c["workload"] and c["value"] > 0 and c["e2e"]["h2d_bytes_per_step"] > 0
    assert all(x["how"].startswith("200 launches in one CUDA graph") for x in d["logit_combine"] if x["B"] <= 64)
    assert d["cpu_baseline"]["kind"] == "port" and "FULL depth" in d["cpu_baseline"]["sample"]
    for n in (2, 4, 8):
        t = json.load(open(os.path.join(ROOT, "profiles", f"r2_bench_tp{n}.json")))
        assert t["n_gpus"] == n and t["scaling"] == "strong" and t["config"]["parallelism"].startswith(f"tp{n}")


def test_round2_reference_arm_fits_its_wall_clock():
    """the reference arm claimed 92.6 s per step inside a 45 s run.  ms_per_step is now the measured wall time of a sample."""
    d = json.load(open(os.path.join(ROOT, "profiles", "r2_bench_reference_arm.json")))
    assert d["impl"] == "reference" and d["value_kind"].startswith("projected")
    assert d["ms_per_step"] * d["steps"] < 120e3 and d["projected_ms_per_full_generate"] > d["ms_per_step"]
    assert "FULL depth" in d["cpu_baseline"]["sample"]
