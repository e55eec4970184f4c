#!/usr/bin/env python
"""Join an `ncu --page source --csv` SASS listing with nvdisasm line info, aggregate per source line.

usage: sass_lines.py <ncu_source.csv> <nvdisasm -g -c output> <kernel name substring> [top]
"""
import csv
import re
import sys
from collections import defaultdict

src_csv, sass, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
# --- nvdisasm: instruction order -> (file, line)
lines = open(sass).read().split("\n")
start = None
for i, l in enumerate(lines):
    if l.startswith(".text.") and kname in l:
        start = i
        break
assert start is not None, "kernel not found"
loc = []
cur = ("?", 0)
for l in lines[start + 1:]:
    if l.startswith(".text.") or l.startswith(".section"):
        if loc:
            break
    m = re.search(r"//## File \"([^\"]+)\", line (\d+)", l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        loc.append(cur)
# --- ncu per-instruction rows
rows = list(csv.reader(open(src_csv)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
h = rows[hi]
ci, si, ti = h.index("Instructions Executed"), h.index("# Samples"), h.index("Thread Instructions Executed")
inst = rows[hi + 1:]
assert abs(len(inst) - len(loc)) < 8, (len(inst), len(loc))
agg = defaultdict(lambda: [0, 0, 0])
for k, r in enumerate(inst):
    if k >= len(loc):
        break
    a = agg[loc[k]]
    a[0] += int(r[ci]); a[1] += int(r[si]); a[2] += int(r[ti])
ti_, ts_ = sum(a[0] for a in agg.values()), sum(a[1] for a in agg.values())
srcs = {}
def text(f, n):
    import glob
    if f not in srcs:
        c = glob.glob(f"/root/repo/rabitq-ann-search_b200/csrc/{f}")
        srcs[f] = open(c[0]).read().split("\n") if c else []
    return srcs[f][n - 1].strip()[:100] if 0 < n <= len(srcs[f]) else ""
print(f"total warp-instructions {ti_}, samples {ts_}")
print("== top by instructions executed")
for (f, n), a in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
    print(f"{a[0] / ti_ * 100:5.1f}% inst {a[1] / ts_ * 100:5.1f}% smp  thr/inst {a[2] / max(a[0], 1):4.1f}  {f}:{n}  {text(f, n)}")
print("== top by stall samples")
for (f, n), a in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
    print(f"{a[0] / ti_ * 100:5.1f}% inst {a[1] / ts_ * 100:5.1f}% smp  thr/inst {a[2] / max(a[0], 1):4.1f}  {f}:{n}  {text(f, n)}")
