#!/usr/bin/env python
"""Generates the committed golden fixtures from the UNMODIFIED reference (oracle/_ref, built from
/root/reference by oracle/Makefile).  Run where /root/reference exists:

    python tests/golden/make_golden.py

Outputs (all small, all deterministic for a given toolchain; g++ 13.3, -O3 -march=x86-64-v3 -mfma):
  k1_golden.npz    queries -> LUT bytes + coefficients from RaBitQEncoder::encode_query_raw
  k2_golden.npz    neighbour blocks + query LUTs -> integer sums and float estimates / bounds from
                   fastscan::compute_* and convert_* (full blocks, partial blocks, degenerate aux values)
  l2_golden.npz    dot_product_simd / l2_distance_simd
  ref_n300_d24_b{1,2,4}.bin   finalized indexes written by CPIndex.save
  e2e_golden.npz   queries -> CPIndex.search_batch ids / distances on those indexes (k = 1, 10, 50)
"""
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parents[1]
sys.path.insert(0, str(ROOT / "oracle"))
sys.path.insert(0, str(ROOT / "tests"))
import cphnsw_oracle as co  # noqa: E402
import common  # noqa: E402


def main():
    assert co.have_ref(), "build oracle/_ref first (make -C oracle ref)"
    o = co.Oracle()
    rng = np.random.default_rng(2024)

    # ---- K1 ----
    k1 = {}
    for dim in (16, 20, 96, 128, 960):
        q = rng.standard_normal((4, dim)).astype(np.float32)
        q[1] *= 100.0
        lut, coeffs = o.ref_encode_queries(q)
        k1[f"q_{dim}"] = q
        k1[f"lut_{dim}"] = lut
        k1[f"coeffs_{dim}"] = coeffs
        k1[f"rot_{dim}"] = o.ref_rotate(q)
    np.savez_compressed(HERE / "k1_golden.npz", **k1)

    # ---- K2 ----
    k2 = {}
    for dim, bits in ((128, 1), (128, 2), (128, 4), (960, 2), (16, 4)):
        fab = common.fabricate(10, dim, bits, seed=dim + bits, counts=(32, 31, 24, 9, 8, 1), degenerate=True, a=1.02, b=0.01)
        q = rng.standard_normal((2, dim)).astype(np.float32)
        lut, coeffs = o.ref_encode_queries(q)
        lay, nb = fab.lay, fab.nb_off
        recs, outs = [], {k: [] for k in ("nbit", "msb", "msb2", "est", "lower", "msb_lower", "dqp", "qi", "count")}
        for v in range(fab.n):
            rec = fab.search_data[v]
            f = lambda name, t, cnt: rec[nb + lay[name]:nb + lay[name] + cnt].view(t)  # noqa: E731
            count = int(f("count", np.uint32, 4)[0])
            qi = v % 2
            dqp = np.float32([37.5, 0.0, 5e-13, 211.25][v % 4])
            nbit, msb, msb2 = o.fastscan(fab.D, bits, lut[qi], rec[nb:nb + 4 * fab.D * bits], ref=True)
            params = np.array([*coeffs[qi], fab.affine_a, fab.affine_b, fab.ip_qo_floor, fab.slack_levels[v % 3]], np.float32)
            wpop = f("wpop", np.uint16, 64) if bits > 1 else None
            est, lower, msb_lower = o.convert(fab.D, bits, params, nbit, msb, msb2, f("nop", np.float32, 128), f("ip_qo", np.float32, 128),
                                              f("ip_cp", np.float32, 128), f("pop", np.uint16, 64), wpop, count, float(dqp), ref=True)
            for k_, val in (("nbit", nbit), ("msb", msb), ("msb2", msb2), ("est", est), ("lower", lower), ("msb_lower", msb_lower)):
                outs[k_].append(val)
            outs["dqp"].append(dqp); outs["qi"].append(qi); outs["count"].append(count)
            recs.append(rec[nb:])
        tag = f"{dim}_{bits}"
        k2[f"blocks_{tag}"] = np.stack(recs)
        k2[f"lut_{tag}"] = lut
        k2[f"coeffs_{tag}"] = coeffs
        k2[f"calib_{tag}"] = np.array([fab.affine_a, fab.affine_b, fab.ip_qo_floor, *fab.slack_levels[:3]], np.float32)
        for k_, val in outs.items():
            k2[f"{k_}_{tag}"] = np.asarray(val)
    np.savez_compressed(HERE / "k2_golden.npz", **k2)

    # ---- exact distances ----
    l2 = {}
    for D in (16, 128, 1024):
        a = rng.standard_normal((6, D)).astype(np.float32)
        b = rng.standard_normal((6, D)).astype(np.float32)
        l2[f"a_{D}"] = a; l2[f"b_{D}"] = b
        l2[f"dot_{D}"] = np.array([o.dot(a[i], b[i], ref=True) for i in range(6)], np.float32)
        l2[f"l2_{D}"] = np.array([o.l2(a[i], b[i], ref=True) for i in range(6)], np.float32)
    np.savez_compressed(HERE / "l2_golden.npz", **l2)

    # ---- end to end ----
    e2e = {}
    base = co.synthetic(300, 24, seed=7)
    q = co.synthetic(40, 24, seed=8)
    e2e["queries"] = q
    for bits in (1, 2, 4):
        path = HERE / f"ref_n300_d24_b{bits}.bin"
        idx = co.build_reference_index(base, bits, path, threads=4)
        for k in (1, 10, 50):
            ids, d = idx.search_batch(q, k)
            e2e[f"ids_b{bits}_k{k}"] = ids
            e2e[f"dists_b{bits}_k{k}"] = d
    np.savez_compressed(HERE / "e2e_golden.npz", **e2e)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
