#!/usr/bin/env python
"""Aggregate an ncu SASS source listing per enclosing source function (and per marked kernel section).

usage: regions.py <ncu_source.csv> <nvdisasm -g -c output> <kernel substring> <expansions>
Kernel sections are delimited in search.cu by comments starting with '// ----'.
"""
import csv, re, sys
from collections import defaultdict
from pathlib import Path

src_csv, sass, kname, nexp = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
CSRC = Path(__file__).resolve().parents[1] / "rabitq-ann-search_b200" / "csrc"

def func_map(path):
    m, cur, depth, name = {}, None, 0, None
    sect = None
    for n, line in enumerate(path.read_text().split("\n"), 1):
        if depth <= 1 and ("__device__" in line or "__global__" in line) and not line.strip().startswith("//"):
            cands = [x for x in re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", line) if not x.startswith("__")]
            if cands:
                name = cands[-1]; sect = None
        if name == "search_kernel" and line.strip().startswith("// ----"):
            sect = line.strip()[7:60].strip(" -")
        m[n] = (name or "?") + ((": " + sect) if (name == "search_kernel" and sect) else "")
        depth += line.count("{") - line.count("}")
        if depth == 0 and "}" in line:
            pass
    return m

maps = {p.name: func_map(p) for p in list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh"))}
lines = open(sass).read().split("\n")
start = [i for i, l in enumerate(lines) if l.startswith(".text.") and kname in l][0]
loc, cur = [], ("?", 0)
for l in lines[start + 1:]:
    if (l.startswith(".text.") or l.startswith(".section")) and loc:
        break
    mm = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if mm:
        cur = (mm.group(1).split("/")[-1], int(mm.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        loc.append(cur)
rows = list(csv.reader(open(src_csv)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
h = rows[hi]; ci, si = h.index("Instructions Executed"), h.index("# Samples")
inst = rows[hi + 1:]
assert abs(len(inst) - len(loc)) < 8, (len(inst), len(loc))
agg = defaultdict(lambda: [0, 0])
for k, r in enumerate(inst[:len(loc)]):
    f, n = loc[k]
    key = maps.get(f, {}).get(n, f)
    agg[key][0] += int(r[ci]); agg[key][1] += int(r[si])
ti, ts = sum(a[0] for a in agg.values()), sum(a[1] for a in agg.values())
print(f"total {ti / nexp:.0f} warp-instructions per expansion")
for k, a in sorted(agg.items(), key=lambda x: -x[1][0]):
    if a[0] / ti > 0.003:
        print(f"{k:62s} {a[0] / ti * 100:5.1f}% inst ({a[0] / nexp:6.1f}/exp) {a[1] / ts * 100:5.1f}% samples")
