"""Runs one small case of an emulated kernel harness built with -fsanitize=thread (tests/test_kernels_racecheck.py starts
this under LD_PRELOAD=libtsan.so and reads the reports from stderr).   python racecheck_driver.py <which> <dir with the libs>"""
import ctypes as C
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
for p in (ROOT / "tests", ROOT / "oracle"):
    sys.path.insert(0, str(p))

import numpy as np  # noqa: E402

import common  # noqa: E402
import test_kernels_emulated as T  # noqa: E402
import test_neighbor_codes_emulated as N  # noqa: E402
from common import co  # noqa: E402

which, libdir = sys.argv[1], Path(sys.argv[2])
lib = lambda name: C.CDLL(str(libdir / f"lib{name}.so"))  # noqa: E731
oracle = co.Oracle()
ok = True
if which == "probe":          # a kernel with a missing / present __syncwarp: the detector must tell them apart
    lib("race_probe").probe(int(sys.argv[3]))
elif which == "k1":
    g = np.load(common.GOLDEN / "k1_golden.npz")
    for dim in (20, 128):
        lut = T._prepare(lib("query_prep_emul"), oracle, g[f"q_{dim}"])[0]
        ok &= np.array_equal(lut, g[f"lut_{dim}"])
    q = np.random.default_rng(1).standard_normal((5, 96)).astype(np.float32)
    T._prepare(lib("query_prep_emul"), oracle, q, center=True, centroid=q[0] * 0.5)
elif which == "k2":
    g = np.load(common.GOLDEN / "k2_golden.npz")
    for tag in ("128_4", "16_4"):
        dim, bits = map(int, tag.split("_"))
        D = max(16, 1 << (dim - 1).bit_length())
        n = g[f"blocks_{tag}"].shape[0]
        outs, _ = T._fastscan(lib("fastscan_emul"), dim, bits, g[f"blocks_{tag}"], g[f"calib_{tag}"][:6], T._uplanes_from_lut(g[f"lut_{tag}"], D),
                              g[f"coeffs_{tag}"], g[f"dqp_{tag}"].astype(np.float32), qi=g[f"qi_{tag}"].astype(np.uint32),
                              levels=(np.arange(n) % 3).astype(np.int32), vertex_ids=np.arange(n, dtype=np.uint32))
        ok &= np.array_equal(outs["nbit"], g[f"nbit_{tag}"])
elif which == "k3":
    g = np.load(common.GOLDEN / "e2e_golden.npz")
    sf = co.SaveFile(common.GOLDEN / "ref_n300_d24_b4.bin")
    ids, d, over, _ = T._search(lib("query_prep_emul"), lib("search_emul"), oracle, sf.dim, 4, sf.search_data, sf.rec_size, sf.nb_off, sf.raw,
                                sf.norm_sq, sf.calib_bytes, sf.max_level, sf.entry_point, sf.layers, g["queries"][:3], 10, warps=4, ctas=2)
    gi, _ = common.sorted_rows(ids, d)
    wi, _ = common.sorted_rows(g["ids_b4_k10"][:3], g["dists_b4_k10"][:3])
    ok &= np.array_equal(gi, wi) and over == 0
elif which == "k5":
    g = np.load(common.GOLDEN / "exhaustive_golden.npz")
    sf = co.SaveFile(common.GOLDEN / "ref_n300_d24_b1.bin")
    q = g["queries"][:2]
    _, _, ids, _ = T._exhaustive(lib("query_prep_emul"), lib("exhaustive_emul"), oracle, sf, q, 10, 100, nslices=2)
    ok &= all(np.array_equal(ids[i], g[f"ids_{i}_k10_kp100"]) for i in range(len(q)))
    # the scan in pieces: candidate mode, warp key selection, CTA select with the earlier pieces' list, merge kernel
    _, _, ids, _ = T._exhaustive(lib("query_prep_emul"), lib("exhaustive_emul"), oracle, sf, q, 10, 100, nslices=2, pieces=3)
    ok &= all(np.array_equal(ids[i], g[f"ids_{i}_k10_kp100"]) for i in range(len(q)))
elif which == "n4":
    g = np.load(common.GOLDEN / "calib_golden.npz")
    sf = co.SaveFile(common.GOLDEN / "ref_n300_d24_b4.bin")
    got = T._calibration(lib("query_prep_emul"), lib("calibration_emul"), oracle, sf.dim, 4, sf.search_data, sf.rec_size, sf.nb_off, sf.raw,
                         g["queries_b4"][:12], g["start_b4"][:12])
    ok &= np.array_equal(got["neighbor"], g["neighbor_b4"][:12]) and np.array_equal(got["true_ip"].view(np.uint32), g["true_ip_b4"][:12].view(np.uint32))
elif which == "n3":
    for dim, bits, global_tile in ((128, 4, False), (96, 2, True), (20, 1, True)):
        vec, pids, nbr = common.neighbor_code_case(dim, 5, dim + bits)
        want, _ = common.expected_neighbor_codes(oracle, dim, bits, vec, pids, nbr)
        codes = N._run(lib("neighbor_codes_emul"), oracle, dim, bits, vec, pids, nbr, want_blocks=True, global_tile=global_tile)[0]
        ok &= np.array_equal(codes, want)
print("RESULTS_OK" if ok else "RESULTS_DIFFER")
