#!/usr/bin/env python
"""Golden vectors for the calibration sample loop (N4): the loop of Index::calibrate_estimator (api/hnsw_index.hpp:786-866)
composed from the UNMODIFIED reference's primitives and block types by oracle/refshim.cpp (refshim_calib_samples), run on the
three committed reference-built index files.  Queries: stored vectors and perturbed stored vectors, as calibrate_estimator
draws them (its RNG is libstdc++'s; the perturbations here are numpy's, which does not matter to the loop).
    python tests/golden/make_calib_golden.py        (needs oracle/_ref: run where /root/reference exists)
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT / "tests"))
import common  # noqa: E402
from common import co  # noqa: E402

o = co.Oracle()
out = {}
for bits in (1, 2, 4):
    sf = co.SaveFile(common.GOLDEN / f"ref_n300_d24_b{bits}.bin")
    rng = np.random.default_rng(40 + bits)
    ns = 96
    start = rng.integers(0, sf.n, ns).astype(np.uint32)
    q = np.ascontiguousarray(sf.raw[rng.integers(0, sf.n, ns), :sf.dim]).astype(np.float32)
    q[ns // 2:] += (0.3 * rng.standard_normal((ns - ns // 2, sf.dim))).astype(np.float32)
    qp = np.zeros((ns, sf.D), np.float32)
    qp[:, :sf.dim] = q
    ref = o.ref_calibration_samples(sf, qp, start)
    out[f"queries_b{bits}"] = q
    out[f"start_b{bits}"] = start
    for k, v in ref.items():
        out[f"{k}_b{bits}"] = v
np.savez_compressed(common.GOLDEN / "calib_golden.npz", **out)
print("wrote", common.GOLDEN / "calib_golden.npz")
