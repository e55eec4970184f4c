#!/usr/bin/env python
"""Turns the ncu artefacts of a round into the committed summaries under profiles/.

usage: summarize.py <launches.csv> <kernels.ncu-rep> <bench.json> <tag>
"""
import csv, json, subprocess, sys
from collections import defaultdict

launches, rep, bench, tag = sys.argv[1:5]
out = []
# ---- launch list: share of the step per kernel
rows = list(csv.reader(open(launches)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hi]; ki, vi = h.index("Kernel Name"), h.index("Metric Value")
d = defaultdict(list)
for r in rows[hi + 1:]:
    if len(r) > vi:
        d[r[ki].split("(")[0]].append(float(r[vi].replace(",", "")))
tot = sum(sum(v) for v in d.values())
out.append(f"# {tag}: launch list (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised)\n")
out.append("| kernel | launches | total us | share |\n|---|---|---|---|\n")
for k, v in sorted(d.items(), key=lambda x: -sum(x[1])):
    out.append(f"| {k} | {len(v)} | {sum(v) / 1e3:.1f} | {sum(v) / tot:.4f} |\n")
# ---- full capture: one row per profiled launch
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.split("\n")))
h = rr[0]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
idx = {w: h.index(w) for w in want if w in h}
out.append(f"\n# {tag}: ncu --set full --clock-control none, per profiled launch (units as ncu prints them: {dict((w, rr[1][i]) for w, i in idx.items() if rr[1][i])})\n\n")
traffic = {}
for r in rr[2:]:
    if len(r) < len(h):
        continue
    name = r[idx["Kernel Name"]].split("(")[0]
    out.append("* " + name + ": " + ", ".join(f"{w.split('.')[0]}={r[i]}" for w, i in idx.items() if w != "Kernel Name") + "\n")
    if "search_kernel" in name and "dram" not in traffic:
        def gb(s, unit):
            v = float(s.replace(",", "")); return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[unit]
        rd = gb(r[idx["dram__bytes_read.sum"]], rr[1][idx["dram__bytes_read.sum"]])
        wr = gb(r[idx["dram__bytes_write.sum"]], rr[1][idx["dram__bytes_write.sum"]])
        traffic = {"dram": rd + wr, "kernel": name, "read": rd, "write": wr}
b = json.loads(open(bench).read().strip().split("\n")[-1])
out.append(f"\n# {tag}: bench line it belongs to\n\n```\n{json.dumps({k: b[k] for k in ('value', 'ms_per_step', 'e2e', 'roofline', 'clocks', 'fastscan_stream', 'cpu_baseline', 'recall_at_10') if k in b}, indent=1)}\n```\n")
open(f"profiles/{tag}_summary.md", "w").write("".join(out))
if traffic:
    json.dump({"dram_bytes_per_launch": traffic["dram"], "read": traffic["read"], "write": traffic["write"], "kernel": traffic["kernel"],
               "source": f"profiles/{tag}_summary.md (ncu --set full, one launch of the 10k-query batch at 1M x 128 x 4-bit)"},
              open("profiles/search_kernel_traffic.json", "w"), indent=1)
print("".join(out))
