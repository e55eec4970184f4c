"""Golden vectors for the exhaustive batched scan (K5), written WITHOUT the C restatement.

    python tests/golden/make_exhaustive_golden.py          # needs oracle/_ref (the compiled, unmodified reference)

The reference has no exhaustive mode; SURVEY section 8c specifies it as a composition of reference
primitives.  This script performs that composition in Python with every primitive executed by the UNMODIFIED reference
through oracle/_ref/libcphnsw_refshim.so -- encode_query_raw (rotation + LUT + coefficients), compute_inner_products on
FastScanCodeBlock<D,32> groups of the per-vertex codes, convert_to_distances_with_bounds (full groups of 32: the AVX2
lanes), dot_product_simd for |qc|^2 and for the exact_l2 lambda -- on the committed index tests/golden/ref_n300_d24_b1.bin,
and stores sums, estimates and search results.  tests/test_oracle_golden.py pins the C restatement
(cpo_exhaustive_search) against the file; tests/test_exhaustive_gpu.py pins the three CUDA forms against it.
"""
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parents[1] / "oracle"))
import cphnsw_oracle as co  # noqa: E402


def compose(o, sf, q, cases):
    dim, D, n = sf.dim, sf.D, sf.n
    storage = -(-(8 * ((D + 63) // 64)) // 64) * 64
    sd = sf.search_data
    code = np.ascontiguousarray(sd[:, :D // 8])                          # signs, bit i of the code = bit i%8 of byte i/8
    nop = np.ascontiguousarray(sd[:, storage:storage + 4]).view(np.float32)[:, 0]
    ipqo = np.ascontiguousarray(sd[:, storage + 4:storage + 8]).view(np.float32)[:, 0]
    pop = np.unpackbits(code, axis=1).sum(1).astype(np.uint16)
    qc = (q - sf.centroid).astype(np.float32)
    lut, coeffs = o.ref_encode_queries(qc[None, :])
    qcp = np.zeros(D, np.float32); qcp[:dim] = qc
    qp = np.zeros(D, np.float32); qp[:dim] = q
    dqp = o.dot(qcp, qcp, ref=True)
    slack = float(sf.slack_levels[0]) if sf.num_slack_levels > 0 else 0.0
    params = np.array([coeffs[0, 0], coeffs[0, 1], coeffs[0, 2], sf.affine_a, sf.affine_b, sf.ip_qo_floor, slack], np.float32)
    sums = np.zeros(n, np.uint32); est = np.zeros(n, np.float32)
    for v0 in range(0, n, 32):
        idx = np.arange(v0, v0 + 32)
        idx[idx >= n] = 0                                                 # pad the last group: every lane on the AVX2 path
        # FastScanCodeBlock<D,32>::store: packed[sp][v] = nibble(seg 2sp+1) << 4 | nibble(seg 2sp) = code byte sp of vertex v
        planes = np.ascontiguousarray(code[idx].T)                        # [D/8][32]
        nbit, _, _ = o.fastscan(D, 1, lut[0], planes, ref=True)
        e, _, _ = o.convert(D, 1, params, nbit, nbit, nbit, nop[idx], ipqo[idx], np.zeros(32, np.float32), pop[idx], None, 32, float(dqp), ref=True)
        m = min(32, n - v0)
        sums[v0:v0 + m] = nbit[:m]; est[v0:v0 + m] = e[:m]
    qn = o.dot(qp, qp, ref=True)
    out = {}
    for k, kp in cases:
        order = np.lexsort((np.arange(n), est.view(np.uint32)))[:min(kp, n)]      # k' smallest (estimate, id)
        d = np.empty(len(order), np.float32)
        for j, vid in enumerate(order):
            dot = o.dot(qp, sf.raw[vid], ref=True)
            r = np.float32(np.float32(qn + sf.norm_sq[vid]) - np.float32(2.0) * dot)
            d[j] = r if r > 0 else np.float32(0.0)
        sel = np.lexsort((order, d.view(np.uint32)))[:k]
        ids = np.full(k, -1, np.int64); dd = np.full(k, np.finfo(np.float32).max, np.float32)
        ids[:len(sel)] = order[sel]; dd[:len(sel)] = d[sel]
        out[(k, kp)] = (ids, dd)
    return sums, est, out


def main():
    assert co.have_ref(), "oracle/_ref is not built"
    o = co.Oracle()
    sf = co.SaveFile(HERE / "ref_n300_d24_b1.bin")
    q = np.load(HERE / "e2e_golden.npz")["queries"][:12].copy()
    q[3] = sf.centroid                                                    # |q - c|^2 == 0: the dist_qp_sq < 1e-12 branch
    cases = [(10, 100), (1, 1), (5, 32), (10, 300)]
    g = {"queries": q}
    for i in range(len(q)):
        sums, est, res = compose(o, sf, q[i], cases)
        g[f"sums_{i}"] = sums; g[f"est_{i}"] = est
        for (k, kp), (ids, dd) in res.items():
            g[f"ids_{i}_k{k}_kp{kp}"] = ids; g[f"dists_{i}_k{k}_kp{kp}"] = dd
    np.savez_compressed(HERE / "exhaustive_golden.npz", **g)
    print("written", HERE / "exhaustive_golden.npz")


if __name__ == "__main__":
    main()
