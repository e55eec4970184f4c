#!/usr/bin/env python
"""How many slots per expansion want their other code planes?  (CPU only; the figure behind search.cu: slot_planes_by_warp.)

Builds a temporary copy of the oracle's C restatement with one counter block in its search loop -- per expansion of a full
result set: the number of NEW slots whose plane-0 lower bound is under the k-th distance -- and runs it over a benchmark
index built by the reference's own code (oracle/_ref/refbuild; `n` vectors x 128, 4-bit, the bench's seed and queries).

    python profiles/candidate_histogram.py [n=250000] [queries=400]

Round 2, n = 250 000: 0: 41.7 %, 1: 32.0 %, 2: 15.5 %, 3: 6.4 %, 4: 2.5 %, >= 5: 2.0 %; some slot (new or not) under the
bound in 98.5 % of the expansions; 11.0 new slots and 4.7 slots under the bound per expansion.
"""
import ctypes as C
import os
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "oracle"))
import cphnsw_oracle as co  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 250_000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 400
tmp = Path(tempfile.mkdtemp(prefix="cand_hist_"))
src = (ROOT / "oracle" / "cphnsw_oracle.c").read_text()
anchor = "        int warmup = nn.n < k;\n        for (uint32_t i = 0; i < nnb; ++i) {"
assert anchor in src, "the oracle's search loop moved: update the anchor"
src = src.replace(anchor, """        int warmup = nn.n < k;
        if (!warmup) {
            extern unsigned long long g_hist[40]; unsigned c = 0, cn = 0, pre = 0;
            float w0 = nn_worst(&nn);
            for (uint32_t i = 0; i < nnb; ++i) { int is_new = !estimated[nb.ids[i]]; if (lower[i] < w0) { ++pre; c += is_new; } cn += is_new; }
            __atomic_fetch_add(&g_hist[c], 1, __ATOMIC_RELAXED); __atomic_fetch_add(&g_hist[36], pre > 0, __ATOMIC_RELAXED);
            __atomic_fetch_add(&g_hist[37], 1, __ATOMIC_RELAXED); __atomic_fetch_add(&g_hist[38], cn, __ATOMIC_RELAXED); __atomic_fetch_add(&g_hist[39], pre, __ATOMIC_RELAXED);
        }
        for (uint32_t i = 0; i < nnb; ++i) {""") + "\nunsigned long long g_hist[40];\n"
(tmp / "cphnsw_oracle.c").write_text(src)
(tmp / "cphnsw_oracle.h").write_text((ROOT / "oracle" / "cphnsw_oracle.h").read_text())
lib = tmp / "liboracle_hist.so"
subprocess.run(["gcc", "-O2", "-std=gnu11", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-fPIC", "-shared", "-mavx2", "-mfma",
                "-o", str(lib), str(tmp / "cphnsw_oracle.c"), "-lm"], check=True)
index = tmp / f"c2_{n}x128_b4.bin"
vec = tmp / "base.f32"
co.synthetic(n, 128, 1234).tofile(vec)
subprocess.run([str(ROOT / "oracle" / "_ref" / "refbuild"), "128", "4", str(n), str(vec), str(index)], check=True,
               env=dict(os.environ, OMP_NUM_THREADS=str(min(16, os.cpu_count() or 1))), stdout=sys.stderr)
vec.unlink()
co.build_port = lambda force=False: lib
o = co.Oracle()
q = np.random.default_rng(99).standard_normal((nq, 128)).astype(np.float32)
_, _, st = o.search_batch(o.index_view(co.SaveFile(index)), q, 10, threads=os.cpu_count() or 1)
h = list((C.c_ulonglong * 40).in_dll(o.lib, "g_hist"))
tot = h[37]
print(f"{n} x 128 x 4-bit, {nq} queries, k = 10: {st['expansions'] / nq:.0f} expansions per query, {tot} with a full result set")
print(f"some slot under the plane-0 bound: {h[36] / tot:.3f}; new slots per expansion {h[38] / tot:.2f}, slots under the bound {h[39] / tot:.2f}")
cum = 0.0
for c in range(33):
    if h[c]:
        cum += h[c] / tot
        print(f"{c:2d} new slots under the bound: {h[c] / tot:.4f}  (cumulative {cum:.4f})")
