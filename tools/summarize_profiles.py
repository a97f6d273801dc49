"""Turns the raw ncu outputs under gpurun_out/ into the tracked summaries under profiles/ (round-tagged).
usage: python tools/summarize_profiles.py r1  [launch_csv] [ncu_rep ...]"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
launch_csv = sys.argv[2]
reps = sys.argv[3:]
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)

# ---- launch list: per-kernel count / total / average device time, and share of one decode step
lines = [l for l in open(launch_csv) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
agg = collections.OrderedDict()
seq = []
for r in rows:
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    name = re.sub(r"^void ", "", name)
    us = float(r["Metric Value"].replace(",", "")) / 1000.0
    seq.append((name, us))
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += us
# one decode step = from one combine_sample to the next
idx = [i for i, (n, _) in enumerate(seq) if "combine_sample" in n]
step = None
if len(idx) >= 3:
    s0, s1 = idx[-2] + 1, idx[-1] + 1
    st = collections.OrderedDict()
    for n, us in seq[s0:s1]:
        a = st.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += us
    tot = sum(v[1] for v in st.values())
    step = {"total_us": tot, "kernels": {k: {"launches": v[0], "total_us": round(v[1], 2), "share": round(v[1] / tot, 4)} for k, v in st.items()}}
out = {"source": f"ncu --metrics gpu__time_duration.sum --clock-control none ... python bench.py --steps 1 --warmup 0 --new-tokens 6 --profile ({os.path.basename(launch_csv)})",
       "note": "per-launch times under ncu are cold-cache and serialised (no PDL overlap): compare SHARES, not absolutes",
       "all_launches": {k: {"launches": v[0], "total_us": round(v[1], 1), "avg_us": round(v[1] / v[0], 2)} for k, v in
                        sorted(agg.items(), key=lambda kv: -kv[1][1]) if not k.startswith("at::")},
       "one_decode_step": step}
json.dump(out, open(os.path.join(ROOT, "profiles", f"{tag}_launch_list.json"), "w"), indent=1)
print("decode step share:", json.dumps(step, indent=1)[:1500] if step else None)

WANT = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__bytes_read.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_op_hmma.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum"]
for rep in reps:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(raw.splitlines()))
    if len(r) < 3:
        continue
    hdr, units = r[0], r[1]
    ks = []
    for row in r[2:]:
        d = {}
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                d[w + (f" [{units[i]}]" if units[i] else "")] = row[i]
        ks.append(d)
    name = os.path.splitext(os.path.basename(rep))[0]
    json.dump({"source": f"ncu --set full --clock-control none --import-source on ({os.path.basename(rep)})", "kernels": ks},
              open(os.path.join(ROOT, "profiles", f"{tag}_{name}.json"), "w"), indent=1)
    # per-launch DRAM traffic of gemv launches
    if "gemv" in name:
        tr = []
        for d in ks:
            rd = [v for k, v in d.items() if k.startswith("dram__bytes_read.sum [")]
            wr = [v for k, v in d.items() if k.startswith("dram__bytes_write.sum [")]
            un = [k for k in d if k.startswith("dram__bytes_read.sum [")][0]
            scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}[un.split("[")[1].rstrip("]")]
            unw = [k for k in d if k.startswith("dram__bytes_write.sum [")][0]
            scalew = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}[unw.split("[")[1].rstrip("]")]
            tr.append(float(rd[0]) * scale + float(wr[0]) * scalew)
        big = [x for x in tr if x > 3e8]  # the fused per-layer chains (405 MB each)
        if big:
            json.dump({"dram_bytes_per_launch": sum(big) / len(big), "launches_sampled": len(big),
                       "what": "dram__bytes_read.sum + dram__bytes_write.sum of the fused per-layer gemv_kernel launches (o+gate_up+down+next qkv)",
                       "algorithmic_bytes_per_launch": 404766720 + 0, "source": os.path.basename(rep)},
                      open(os.path.join(ROOT, "profiles", "gemv_traffic.json"), "w"), indent=1)
        print(name, "dram bytes per launch:", [round(x / 1e6, 1) for x in tr])
