#!/usr/bin/env python
"""Instruction histogram per kernel of the shipped library (cuobjdump -sass): the mnemonics that show what the kernels are
made of -- bulk asynchronous copies (UBLKCP), tcgen05 (UTC*MMA, LDTM, UTCBAR), mbarriers (SYNCS), popcounts, logic,
atomics.  usage: sass_summary.py [library.so] > profiles/sass_rNN.md"""
import re
import subprocess
import sys
from collections import Counter, OrderedDict
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
lib = sys.argv[1] if len(sys.argv) > 1 else str(ROOT / "rabitq-ann-search_b200" / "cphnsw_b200" / "libcphnsw_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
WATCH = ["UBLKCP", "UTMALDG", "UTCHMMA", "UTCIMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UTCATOM", "SYNCS", "POPC", "LOP3", "FFMA", "MUFU", "SHFL", "VOTE", "MATCH",
         "ATOMG", "ATOM", "RED", "LDG", "STG", "LDS", "STS", "LD", "ST", "LDL", "STL", "BAR", "HMMA", "IMMA"]
kern = OrderedDict()
cur = None
for line in txt.split("\n"):
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        kern[cur] = Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        op = m.group(1)
        kern[cur][op] += 1
        kern[cur]["_total"] += 1
demangle = subprocess.run(["c++filt"], input="\n".join(kern), capture_output=True, text=True).stdout.split("\n")
print(f"# SASS instruction histogram per kernel of {Path(lib).name} (static counts; `cuobjdump -sass`, sm_100a)\n")
print("| kernel | instructions | " + " | ".join(WATCH) + " |")
print("|---|---|" + "---|" * len(WATCH))
for (name, c), dm in zip(kern.items(), demangle):
    short = re.sub(r"\(.*", "", dm).replace("void ", "").replace("cpb::", "")
    cells = []
    for w in WATCH:
        n = sum(v for k, v in c.items() if k == w or (w in ("LD", "ST") and k in (w + ".E",)) )
        cells.append(str(n) if n else "")
    if c["_total"] < 40:
        continue
    print(f"| `{short}` | {c['_total']} | " + " | ".join(cells) + " |")
