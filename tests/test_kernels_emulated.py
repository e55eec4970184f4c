"""Kernel source on the CPU: the PTX-free kernels of rabitq-ann-search_b200/csrc compiled for the host over
tests/native/cuda_emul.h (a thread per lane, barriers for the warp primitives, IEEE operations for the _rn intrinsics)
and held to the reference's committed golden vectors.  The GPU runs of the same comparisons are in tests/test_parity_gpu.py;
the build-side encoder has its own file (tests/test_neighbor_codes_emulated.py)."""
import ctypes as C
import subprocess

import numpy as np
import pytest

import os

import common

NATIVE = common.ROOT / "tests" / "native"


def _build(tmp_path_factory, name):
    out = tmp_path_factory.mktemp("emul") / f"lib{name}.so"
    # CPHNSW_EMUL_DEFINES: the -D flags of a kernel variant under test (build.py --variant), space separated
    defines = [f"-D{d}" for d in os.environ.get("CPHNSW_EMUL_DEFINES", "").split()]
    cmd = ["g++", "-std=c++20", "-O1", "-ffp-contract=off", "-fPIC", "-shared", "-pthread", "-I/usr/local/cuda/include", *defines,
           str(NATIVE / f"{name}.cpp"), "-o", str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return C.CDLL(str(out))


@pytest.fixture(scope="module")
def k1(tmp_path_factory):
    return _build(tmp_path_factory, "query_prep_emul")


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def _prepare(k1, oracle, q, center=False, centroid=None):
    nq, dim = q.shape
    D = max(16, 1 << (dim - 1).bit_length())
    W = max(D, 128) // 32
    signs = oracle.rotation_signs(D)
    cen = np.zeros(dim, np.float32) if centroid is None else np.ascontiguousarray(centroid, np.float32)
    lut = np.full((nq, D // 4, 16), 0xEE, np.uint8)
    coeffs = np.full((nq, 8), np.nan, np.float32)
    rot = np.full((nq, D), np.nan, np.float32)
    upl = np.full((nq, 4, W), 0xDDDDDDDD, np.uint32)
    qT = np.full((nq, D), np.nan, np.float32)
    q = np.ascontiguousarray(q, np.float32)
    rc = k1.emul_prepare_queries(C.c_uint32(dim), _p(signs, C.c_float), _p(cen, C.c_float), _p(q, C.c_float), C.c_uint32(nq),
                                 C.c_int(int(center)), _p(lut, C.c_uint8), _p(coeffs, C.c_float), _p(rot, C.c_float),
                                 _p(upl, C.c_uint32), _p(qT, C.c_float))
    assert rc == 0
    return lut, coeffs, rot, upl, qT


@pytest.mark.parametrize("dim", [16, 20, 96, 128, 960])
def test_k1_source_reproduces_the_reference_golden_vectors(k1, oracle, dim):
    g = np.load(common.GOLDEN / "k1_golden.npz")
    q = g[f"q_{dim}"]
    lut, coeffs, rot, upl, qT = _prepare(k1, oracle, q)
    assert np.array_equal(lut, g[f"lut_{dim}"])
    assert np.array_equal(_bits(coeffs[:, :3]), _bits(g[f"coeffs_{dim}"]))
    assert np.array_equal(_bits(rot), _bits(g[f"rot_{dim}"]))
    # the bit-plane form of the same 4-bit values (what K2/K3/K5 consume): u_i = lut[i/4][1 << i%4]
    D = rot.shape[1]
    u = np.stack([g[f"lut_{dim}"][:, :, 1 << b] for b in range(4)], 2).reshape(len(q), D).astype(np.uint32)
    W = upl.shape[2]
    for t in range(4):
        bits = np.zeros((len(q), W * 32), np.uint32)
        bits[:, :D] = (u >> t) & 1
        words = (bits.reshape(len(q), W, 32) << np.arange(32, dtype=np.uint32)).sum(2).astype(np.uint32)
        assert np.array_equal(upl[:, t], words), t
    # |q|^2 by dot_product_simd's eight chains, and the accumulator-major copy of the padded query
    for i in range(len(q)):
        pad = np.zeros(D, np.float32); pad[:dim] = q[i]
        assert _bits(coeffs[i, 3:4])[0] == _bits(np.float32(oracle.dot(pad, pad)))[0]
        assert np.array_equal(qT[i].reshape(8, D // 8), pad.reshape(D // 8, 8).T)


@pytest.mark.parametrize("dim", [20, 96, 300])
def test_k1_source_centred_queries_equal_the_oracle(k1, oracle, dim):
    """center=True (the exhaustive scan's query form): q - centroid before the rotation."""
    rng = np.random.default_rng(dim)
    q = rng.standard_normal((9, dim)).astype(np.float32)
    cen = rng.standard_normal(dim).astype(np.float32) * 0.3
    q[2] = cen
    lut, coeffs, rot, _, _ = _prepare(k1, oracle, q, center=True, centroid=cen)
    olut, oco, orot = oracle.encode_queries((q - cen).astype(np.float32), want_rotated=True)
    assert np.array_equal(lut, olut)
    assert np.array_equal(_bits(coeffs[:, :3]), _bits(oco))
    assert np.array_equal(_bits(rot), _bits(orot))


# ---- index hand-off (relayout.cu) + K2 (fastscan_blocks.cu) ------------------------------------------------------------
@pytest.fixture(scope="module")
def k2(tmp_path_factory):
    return _build(tmp_path_factory, "fastscan_emul")


def _uplanes_from_lut(lut, D):
    nq = lut.shape[0]
    u = np.stack([lut[:, :, 1 << b] for b in range(4)], axis=2).reshape(nq, D)
    W = max(D, 128) // 32
    up = np.zeros((nq, 4, W), np.uint32)
    for t in range(4):
        pad = np.zeros((nq, W * 32), np.uint8)
        pad[:, :D] = (u >> t) & 1
        up[:, t, :] = np.packbits(pad.reshape(nq, W, 32), axis=2, bitorder="little").view(np.uint32)[:, :, 0]
    return up


def _fastscan(k2, dim, bits, blocks, calib, up, coeffs, dqp, qi=None, levels=None, vertex_ids=None, lean=False):
    n = blocks.shape[0]
    nblocks = n if vertex_ids is None else len(vertex_ids)
    blocks = np.ascontiguousarray(blocks)
    co8 = np.zeros((up.shape[0], 8), np.float32)
    co8[:, :3] = coeffs
    outs = {name: np.full((nblocks, 32), 0xABABABAB, np.uint32) for name in ("nbit", "msb", "msb2")}
    outs.update({name: np.full((nblocks, 32), np.nan, np.float32) for name in ("est", "lower", "msb_lower")})
    problems = np.zeros(2, np.uint32)
    calib = np.ascontiguousarray(calib, np.float32)
    opt = lambda a, t: None if a is None else _p(np.ascontiguousarray(a), t)  # noqa: E731
    keep = [np.ascontiguousarray(a) if a is not None else None for a in (qi, vertex_ids, levels)]
    want = (lambda name: None) if lean else (lambda name: _p(outs[name], C.c_uint32 if outs[name].dtype == np.uint32 else C.c_float))
    rc = k2.emul_fastscan_blocks(
        C.c_uint32(dim), C.c_uint32(bits), _p(blocks, C.c_uint8), C.c_uint64(blocks.shape[1]), C.c_uint64(n), _p(calib, C.c_float),
        C.c_int(len(calib) - 3), _p(up, C.c_uint32), _p(co8, C.c_float), C.c_uint32(up.shape[0]),
        None if keep[0] is None else _p(keep[0], C.c_uint32), None if keep[1] is None else _p(keep[1], C.c_uint32), C.c_uint64(nblocks),
        _p(dqp, C.c_float), None if keep[2] is None else _p(keep[2], C.c_int32),
        want("nbit"), want("msb"), want("msb2"), _p(outs["est"], C.c_float), _p(outs["lower"], C.c_float), want("msb_lower"),
        C.c_int(int(lean)), _p(problems, C.c_uint32))
    assert rc == 0, rc
    del opt
    return outs, problems


@pytest.mark.parametrize("tag", ["128_1", "128_2", "128_4", "960_2", "16_4"])
def test_relayout_and_k2_sources_reproduce_the_reference_golden_vectors(k2, tag):
    """Reference neighbour blocks -> the product's re-layout kernel -> the product's FastScan kernel (general form: block
    list, per-block query and slack level) = the reference's integer sums and float estimates / bounds, bit for bit."""
    g = np.load(common.GOLDEN / "k2_golden.npz")
    dim, bits = map(int, tag.split("_"))
    D = max(16, 1 << (dim - 1).bit_length())
    blocks = g[f"blocks_{tag}"]
    n = blocks.shape[0]
    up = _uplanes_from_lut(g[f"lut_{tag}"], D)
    outs, problems = _fastscan(k2, dim, bits, blocks, g[f"calib_{tag}"][:6], up, g[f"coeffs_{tag}"], g[f"dqp_{tag}"].astype(np.float32),
                               qi=g[f"qi_{tag}"].astype(np.uint32), levels=(np.arange(n) % 3).astype(np.int32),
                               vertex_ids=np.arange(n, dtype=np.uint32))
    for v in range(n):
        c = int(g[f"count_{tag}"][v])
        for name in ("nbit", "msb", "msb2"):
            assert np.array_equal(outs[name][v], g[f"{name}_{tag}"][v]), (name, v)
        for name in ("est", "lower", "msb_lower"):
            assert np.array_equal(_bits(outs[name][v][:c]), _bits(g[f"{name}_{tag}"][v][:c])), (name, v)
            assert (outs[name][v][c:] == np.finfo(np.float32).max).all()


@pytest.mark.parametrize("dim,bits", [(128, 4), (96, 1), (300, 2)])
def test_k2_streaming_specialisation_equals_the_general_form(k2, dim, bits):
    """LEAN (contiguous range, one query, slack level 0, est + lower only: the kernel behind the HBM figure)."""
    fab = common.fabricate(37, dim, bits, seed=dim, counts=(32, 31, 9, 1), degenerate=True, a=1.01, b=0.002)
    blocks = fab.search_data[:, fab.nb_off:]
    rng = np.random.default_rng(3)
    up = rng.integers(0, 2 ** 32, (1, 4, max(fab.D, 128) // 32), dtype=np.uint64).astype(np.uint32)
    if fab.D < 128:
        up[:, :, fab.D // 32:] = 0
    coeffs = np.array([[0.013, -0.4, 0.07]], np.float32)
    dqp = rng.uniform(0.5, 40.0, fab.n).astype(np.float32)
    import struct
    cal = np.array(struct.unpack_from("<3f", fab.calibration, 0) + struct.unpack_from("<3f", fab.calibration, 108), np.float32)
    general, _ = _fastscan(k2, dim, bits, blocks, cal, up, coeffs, dqp)
    lean, _ = _fastscan(k2, dim, bits, blocks, cal, up, coeffs, dqp, lean=True)
    assert np.array_equal(_bits(lean["est"]), _bits(general["est"]))
    assert np.array_equal(_bits(lean["lower"]), _bits(general["lower"]))


# ---- K3 (+ fused K4): search.cu, over the re-laid-out index and K1's prepared queries --------------------------------
@pytest.fixture(scope="module")
def k3(tmp_path_factory):
    return _build(tmp_path_factory, "search_emul")


def _level_csr(layers, max_level, entry_point):
    """The slot-addressed CSR cphnsw_b200_upload derives from the save file's sorted edge lists (find_edge becomes a
    table: neighbour -> its slot in the same level, node -> its slot one level down)."""
    out, sizes, entry_slot = [], [], 0xFFFFFFFF
    n_levels = len(layers) if max_level > 0 else 0
    for L in range(n_levels):
        nodes, offs, nbrs = (np.ascontiguousarray(a, np.uint32) for a in layers[L])

        def slot_of(arr, ids):
            pos = np.searchsorted(arr, ids)
            ok = (pos < len(arr)) & (arr[np.minimum(pos, len(arr) - 1)] == ids)
            return np.where(ok, pos, 0xFFFFFFFF).astype(np.uint32)

        nbr_slot = slot_of(nodes, nbrs)
        down = nodes.copy() if L == 0 else slot_of(np.ascontiguousarray(layers[L - 1][0], np.uint32), nodes)
        out += [nodes, offs, nbrs, nbr_slot, down]
        sizes.append(len(nodes))
        if L + 1 == max_level:
            entry_slot = int(slot_of(nodes, np.array([entry_point], np.uint32))[0])
    return out, np.array(sizes + [0], np.uint32), entry_slot, n_levels


def _search(k1, k3, oracle, dim, bits, records, rec_size, nb_off, raw, norm_sq, calib, max_level, entry_point, layers, queries, k,
            warps=4, ctas=2, beam_capacity=0, want_stats=False):
    n = raw.shape[0]
    _, coeffs, _, upl, qT = _prepare(k1, oracle, queries)
    arrays, sizes, entry_slot, n_levels = _level_csr(layers, max_level, entry_point)
    ptrs = (C.POINTER(C.c_uint32) * max(1, len(arrays)))(*[_p(a, C.c_uint32) for a in arrays])
    nq = len(queries)
    ids = np.full((nq, k), -77, np.int64)
    dists = np.full((nq, k), np.nan, np.float32)
    stats = np.zeros(11, np.uint64)
    over = C.c_uint32(0)
    records = np.ascontiguousarray(records)
    raw = np.ascontiguousarray(raw, np.float32)
    norm_sq = np.ascontiguousarray(norm_sq, np.float32)
    cal = np.frombuffer(bytes(calib), np.uint8).copy()
    rc = k3.emul_search(C.c_uint32(dim), C.c_uint32(bits), _p(records, C.c_uint8), C.c_uint64(rec_size), C.c_uint32(nb_off), C.c_uint64(n),
                        _p(raw, C.c_float), _p(norm_sq, C.c_float), _p(cal, C.c_uint8), C.c_int32(max_level), C.c_uint32(entry_point),
                        C.c_uint32(entry_slot), C.c_uint32(entry_point), C.c_uint32(n_levels), ptrs, _p(sizes, C.c_uint32),
                        _p(qT, C.c_float), _p(upl, C.c_uint32), _p(coeffs, C.c_float), C.c_uint32(nq), C.c_uint32(k),
                        _p(ids, C.c_int64), _p(dists, C.c_float), C.c_int(warps), C.c_int(ctas), C.c_uint32(beam_capacity),
                        _p(stats, C.c_uint64) if want_stats else None, C.byref(over))
    assert rc == 0, rc
    return ids, dists, over.value, stats


@pytest.mark.parametrize("bits,k,warps", [(1, 10, 2), (2, 50, 4), (4, 10, 1), (4, 1, 4)])
def test_search_kernel_source_reproduces_the_reference_results(k1, k3, oracle, bits, k, warps):
    """Index files built and saved by the unmodified reference, and its own search_batch results on them (e2e_golden):
    re-layout kernels + K1 + the search kernel, all from the product's source, on host threads -- ids and distance bits."""
    g = np.load(common.GOLDEN / "e2e_golden.npz")
    sf = co_SaveFile(common.GOLDEN / f"ref_n300_d24_b{bits}.bin")
    ids, dists, over, _ = _search(k1, k3, oracle, sf.dim, bits, sf.search_data, sf.rec_size, sf.nb_off, sf.raw, sf.norm_sq, sf.calib_bytes,
                                  sf.max_level, sf.entry_point, sf.layers, g["queries"], k, warps=warps, ctas=2 * (4 // warps))   # warps = 1: the single-warp-CTA build
    assert over == 0
    gi, gd = common.sorted_rows(ids, dists)
    wi, wd = common.sorted_rows(g[f"ids_b{bits}_k{k}"], g[f"dists_b{bits}_k{k}"])
    assert np.array_equal(gi, wi) and np.array_equal(_bits(gd), _bits(wd))


def co_SaveFile(path):
    from common import co

    return co.SaveFile(path)


_STAT_FIELDS = ("pops", "expansions", "exact_calls", "beam_pushes", "max_beam", "nn_pushes", "lb_skips", "gamma_terms", "msb_skipped",
                "estimated", "descent_dists")


@pytest.mark.parametrize("dim,bits,k", [(128, 4, 10), (128, 1, 3), (96, 2, 33), (20, 4, 130)])
def test_search_kernel_source_equals_the_oracle_on_fabricated_indexes(k1, k3, oracle, dim, bits, k):
    """Random 'finalized indexes' (partial and empty blocks, degenerate aux values, two upper layers; small gamma and wide
    slack so that gamma termination, lower-bound skips and MSB-only skips fire): results and the kernel's counters against
    the C restatement of rabitq_search::search.  D = 128 takes the compile-time-dimension kernel."""
    fab = common.fabricate(400, dim, bits, seed=dim + bits, counts=(32, 32, 31, 24, 9, 0), degenerate=True, layers=2,
                           gamma=1.02, gamma_max=1.6, gamma_beta=0.8, gamma_warmup=4, floor=0.35, slacks=(0.05, 0.08, 0.1, 0.12))
    q = np.random.default_rng(11).standard_normal((12, dim)).astype(np.float32)
    q[0] = fab.raw[3, :dim]
    ids, dists, over, st = _search(k1, k3, oracle, dim, bits, fab.search_data, fab.search_data.shape[1], fab.nb_off, fab.raw, fab.norm_sq,
                                   fab.calibration, fab.max_level, fab.entry_point, fab.layers, q, k, want_stats=True)
    assert over == 0
    oid, od, ost = oracle.search_batch(oracle.index_view(fab), q, k)
    gi, gd = common.sorted_rows(ids, dists)
    wi, wd = common.sorted_rows(oid, od)
    assert np.array_equal(gi, wi) and np.array_equal(_bits(gd), _bits(wd))
    got = dict(zip(_STAT_FIELDS, (int(v) for v in st)))
    for key in ("pops", "expansions", "beam_pushes", "nn_pushes", "lb_skips", "gamma_terms", "msb_skipped", "estimated", "descent_dists",
                "max_beam"):
        assert got[key] == ost[key], (key, got[key], ost[key])
    assert got["exact_calls"] >= ost["exact_calls"]


def test_search_kernel_source_reports_frontier_overflow(k1, k3, oracle):
    """A frontier arena too small for the query: the kernel must list the query for the re-run instead of writing past it."""
    fab = common.fabricate(400, 64, 1, seed=5, layers=1, gamma=1e6, gamma_max=1e7)
    q = np.random.default_rng(2).standard_normal((4, 64)).astype(np.float32)
    _, _, over, _ = _search(k1, k3, oracle, 64, 1, fab.search_data, fab.search_data.shape[1], fab.nb_off, fab.raw, fab.norm_sq,
                            fab.calibration, fab.max_level, fab.entry_point, fab.layers, q, 5, beam_capacity=64)
    assert over > 0


# ---- K5, popcount form (exhaustive.cu): scan + select / exact re-rank -------------------------------------------------
@pytest.fixture(scope="module")
def k5(tmp_path_factory):
    return _build(tmp_path_factory, "exhaustive_emul")


def _exhaustive(k1, k5, oracle, sf, queries, k, kprime, nslices=1, dense=False, id_begin=0, id_end=None, pieces=1):
    id_end = sf.n if id_end is None else id_end
    _, coeffs, _, upl, qT = _prepare(k1, oracle, queries, center=True, centroid=sf.centroid)
    nq, m = len(queries), id_end - id_begin
    sums = np.full((nq, m), 0xABABABAB, np.uint32) if dense else None
    est = np.full((nq, m), np.nan, np.float32) if dense else None
    ids = np.full((nq, max(k, 1)), -77, np.int64)
    dists = np.full((nq, max(k, 1)), np.nan, np.float32)
    records = np.ascontiguousarray(sf.search_data)
    raw = np.ascontiguousarray(sf.raw, np.float32)
    norm_sq = np.ascontiguousarray(sf.norm_sq, np.float32)
    cal = np.array([sf.affine_a, sf.affine_b, sf.ip_qo_floor], np.float32)
    rc = k5.emul_exhaustive(C.c_uint32(sf.dim), _p(records, C.c_uint8), C.c_uint64(sf.rec_size), C.c_uint32(sf.nb_off), C.c_uint64(sf.n),
                            _p(raw, C.c_float), _p(norm_sq, C.c_float), _p(cal, C.c_float), _p(qT, C.c_float), _p(upl, C.c_uint32),
                            _p(coeffs, C.c_float), C.c_uint32(nq), C.c_uint64(id_begin), C.c_uint64(id_end), C.c_uint32(k),
                            C.c_uint32(kprime), C.c_uint32(nslices), None if sums is None else _p(sums, C.c_uint32),
                            None if est is None else _p(est, C.c_float), _p(ids, C.c_int64), _p(dists, C.c_float), C.c_uint32(pieces))
    assert rc == 0, rc
    return sums, est, ids, dists


def test_exhaustive_kernel_sources_reproduce_the_reference_composition(k1, k5, oracle):
    """tests/golden/exhaustive_golden.npz was composed from primitives executed by the unmodified reference on its own
    1-bit index file: integer sums, estimate bits, and (k, k') results -- here from the scan and select/re-rank kernels'
    source on host threads, with one and with several vertex slices per query tile."""
    g = np.load(common.GOLDEN / "exhaustive_golden.npz")
    sf = co_SaveFile(common.GOLDEN / "ref_n300_d24_b1.bin")
    q = g["queries"]
    sums, est, _, _ = _exhaustive(k1, k5, oracle, sf, q, 0, 0, dense=True)
    for i in range(len(q)):
        assert np.array_equal(sums[i], g[f"sums_{i}"]), i
        assert np.array_equal(_bits(est[i]), _bits(g[f"est_{i}"])), i
    # one pass, and the scan in pieces (candidate mode, thresholds handed on, warp key selection / CTA select, merge kernel)
    for k, kp, nslices, pieces in ((10, 100, 1, 1), (10, 100, 3, 1), (1, 1, 1, 1), (5, 32, 3, 1), (10, 300, 1, 1),
                                   (10, 100, 2, 3), (5, 32, 1, 4), (10, 300, 2, 2), (1, 1, 1, 3)):
        _, _, ids, dists = _exhaustive(k1, k5, oracle, sf, q, k, kp, nslices=nslices, pieces=pieces)
        for i in range(len(q)):
            assert np.array_equal(ids[i], g[f"ids_{i}_k{k}_kp{kp}"]), (i, k, kp, nslices, pieces)
            assert np.array_equal(_bits(dists[i]), _bits(g[f"dists_{i}_k{k}_kp{kp}"])), (i, k, kp, nslices, pieces)


@pytest.mark.parametrize("D", [16, 128, 1024])
def test_exact_distance_kernel_source_against_the_reference_dot_products(k1, k3, oracle, D):
    """K4 primitive: max(|q|^2 + |x|^2 - 2 dot(q, x), 0) (search/rabitq_search.hpp:88-93) with the reference's own
    dot_product_simd values (l2_golden.npz) -- the eight accumulator chains and their reduction order."""
    g = np.load(common.GOLDEN / "l2_golden.npz")
    a, b = g[f"a_{D}"], g[f"b_{D}"]
    _, coeffs, _, _, qT = _prepare(k1, oracle, a)
    norm_sq = np.array([oracle.dot(x, x) for x in b], np.float32)
    ids = np.tile(np.arange(6, dtype=np.uint32), (6, 1))
    out = np.full((6, 6), np.nan, np.float32)
    raw = np.ascontiguousarray(b, np.float32)
    assert k3.emul_exact_l2(C.c_uint32(D), C.c_uint64(6), _p(raw, C.c_float), _p(norm_sq, C.c_float), _p(qT, C.c_float),
                            _p(coeffs, C.c_float), C.c_uint32(6), _p(ids, C.c_uint32), C.c_uint32(6), _p(out, C.c_float)) == 0
    for i in range(6):
        qn = np.float32(oracle.dot(a[i], a[i]))
        want = np.float32(np.float32(qn + norm_sq[i]) - np.float32(np.float32(2.0) * g[f"dot_{D}"][i]))
        want = np.float32(0.0) if want < 0 else want
        assert _bits(out[i, i:i + 1])[0] == _bits(np.array([want]))[0], i
        for j in range(6):   # the other pairs against the oracle's dot
            w = np.float32(np.float32(qn + norm_sq[j]) - np.float32(np.float32(2.0) * oracle.dot(a[i], b[j])))
            assert _bits(out[i, j:j + 1])[0] == _bits(np.array([max(w, np.float32(0.0))]))[0], (i, j)


# ---- N4 (calibration.cu): the sample loop of calibrate_estimator -------------------------------------------------------
@pytest.fixture(scope="module")
def n4(tmp_path_factory):
    return _build(tmp_path_factory, "calibration_emul")


def _calibration(k1, n4, oracle, dim, bits, records, rec_size, nb_off, raw, queries, start):
    ns = len(queries)
    _, coeffs, _, upl, qT = _prepare(k1, oracle, queries)
    records = np.ascontiguousarray(records)
    raw = np.ascontiguousarray(raw, np.float32)
    q = np.ascontiguousarray(queries, np.float32)
    st = np.ascontiguousarray(start, np.uint32)
    o = {"parent": np.full(ns, 0xEEEEEEEE, np.uint32), "nn_dist_sq": np.full(ns, np.nan, np.float32), "dist_qp_sq": np.full(ns, np.nan, np.float32),
         "neighbor": np.full((ns, 32), 0xEEEEEEEE, np.uint32)}
    for name in ("nop", "ip_corrected", "ip_qo_denom", "true_ip"):
        o[name] = np.full((ns, 32), np.nan, np.float32)
    rc = n4.emul_calibration_samples(C.c_uint32(dim), C.c_uint32(bits), _p(records, C.c_uint8), C.c_uint64(rec_size), C.c_uint32(nb_off),
                                     C.c_uint64(raw.shape[0]), _p(raw, C.c_float), _p(q, C.c_float), _p(st, C.c_uint32), C.c_uint64(ns),
                                     _p(qT, C.c_float), _p(upl, C.c_uint32), _p(coeffs, C.c_float), _p(o["parent"], C.c_uint32),
                                     _p(o["nn_dist_sq"], C.c_float), _p(o["dist_qp_sq"], C.c_float), _p(o["nop"], C.c_float),
                                     _p(o["ip_corrected"], C.c_float), _p(o["ip_qo_denom"], C.c_float), _p(o["true_ip"], C.c_float),
                                     _p(o["neighbor"], C.c_uint32))
    assert rc == 0, rc
    return o


@pytest.mark.parametrize("bits", [1, 2, 4])
def test_calibration_kernel_source_on_the_reference_index_files(k1, n4, oracle, bits):
    """calibration.cu on host threads against the composition over the unmodified reference (calib_golden.npz) on the index
    files the reference built and saved: everything bit for bit, except ip_corrected, which carries coeff_constant -- fused
    one way in the search path that K1 reproduces and, in the composition's D = 32 instantiation only, the other way (an
    ulp of a term ~20; tests/test_oracle_calibration.py)."""
    g = np.load(common.GOLDEN / "calib_golden.npz")
    sf = co_SaveFile(common.GOLDEN / f"ref_n300_d24_b{bits}.bin")
    got = _calibration(k1, n4, oracle, sf.dim, bits, sf.search_data, sf.rec_size, sf.nb_off, sf.raw, g[f"queries_b{bits}"], g[f"start_b{bits}"])
    for k in ("parent", "neighbor"):
        assert np.array_equal(got[k], g[f"{k}_b{bits}"]), k
    for k in ("nn_dist_sq", "dist_qp_sq", "nop", "ip_qo_denom", "true_ip"):
        assert np.array_equal(_bits(got[k]), _bits(g[f"{k}_b{bits}"])), k
    assert np.allclose(got["ip_corrected"], g[f"ip_corrected_b{bits}"], rtol=0, atol=2e-5)


@pytest.mark.parametrize("dim,bits", [(128, 4), (100, 2), (20, 1)])
def test_calibration_kernel_source_equals_the_restatement(k1, n4, oracle, dim, bits):
    """Fabricated indexes with partial and empty blocks: every field bit for bit against cpo_calibration_sample."""
    fab = common.fabricate(300, dim, bits, seed=dim * 3 + bits, counts=(32, 32, 29, 9, 0))
    rng = np.random.default_rng(8)
    ns = 40
    start = rng.integers(0, fab.n, ns).astype(np.uint32)
    q = np.ascontiguousarray(fab.raw[rng.integers(0, fab.n, ns), :dim])
    q[ns // 2:] += (0.2 * rng.standard_normal((ns - ns // 2, dim))).astype(np.float32)
    got = _calibration(k1, n4, oracle, dim, bits, fab.search_data, fab.search_data.shape[1], fab.nb_off, fab.raw, q, start)
    qp = np.zeros((ns, fab.D), np.float32)
    qp[:, :dim] = q
    want = oracle.calibration_samples(oracle.index_view(fab), qp, start)
    for k in ("parent", "neighbor"):
        assert np.array_equal(got[k], want[k]), k
    for k in ("nn_dist_sq", "dist_qp_sq", "nop", "ip_corrected", "ip_qo_denom", "true_ip"):
        assert np.array_equal(_bits(got[k]), _bits(want[k])), k


@pytest.mark.parametrize("dim,bits,k", [(128, 4, 10), (300, 4, 10), (960, 2, 20), (200, 2, 5)])
def test_fast_search_kernel_source_equals_the_oracle_on_fabricated_indexes(k1, k3, oracle, dim, bits, k):
    """The non-counting instantiations (result list in registers; the other planes of the few slots that want them evaluated by
    the whole warp, in groups of 16 lanes at D = 128 / 256 and by the warp beyond) against the C restatement: ids and distance bits."""
    fab = common.fabricate(300, dim, bits, seed=2 * dim + bits, counts=(32, 32, 30, 17, 8, 0), degenerate=True, layers=1,
                           gamma=1.05, gamma_max=1.8, gamma_beta=0.8, gamma_warmup=4, floor=0.35, slacks=(0.05, 0.08, 0.1))
    q = np.random.default_rng(5).standard_normal((6, dim)).astype(np.float32)
    ids, dists, over, _ = _search(k1, k3, oracle, dim, bits, fab.search_data, fab.search_data.shape[1], fab.nb_off, fab.raw, fab.norm_sq,
                                  fab.calibration, fab.max_level, fab.entry_point, fab.layers, q, k)
    assert over == 0
    oid, od, _ = oracle.search_batch(oracle.index_view(fab), q, k)
    gi, gd = common.sorted_rows(ids, dists)
    wi, wd = common.sorted_rows(oid, od)
    assert np.array_equal(gi, wi) and np.array_equal(_bits(gd), _bits(wd))


@pytest.mark.parametrize("bits,nch", [(4, 1), (2, 1), (2, 2), (4, 2), (2, 8), (4, 8), (4, 16)])
def test_slot_planes_by_warp_equals_the_lane_per_slot_sums(k3, bits, nch):
    """The warp-cooperative evaluation of a few slots' other planes (search.cu: slot_planes_by_warp; 16-lane groups up to 16
    words per slot, the whole warp beyond) gives the integers plane_sum_one gives, for every choice of slots."""
    rng = np.random.default_rng(100 * bits + nch)
    planes = rng.integers(0, 256, bits * nch * 32 * 16, dtype=np.uint8)
    uq = rng.integers(0, 2**32, 4 * nch * 4, dtype=np.uint32)
    for slots in (0x1, 0x80000000, 0x00010002, 0x80000001, 0x00F00000, 0x12345678, 0xFFFFFFFF, 0x0000A005, 0x40000000):
        out = [np.zeros(32, np.uint32) for _ in range(4)]
        rc = k3.emul_slot_planes(C.c_uint32(bits), C.c_uint32(nch), _p(planes, C.c_uint8), _p(uq, C.c_uint32), C.c_uint32(slots),
                                 *[_p(o, C.c_uint32) for o in out])
        assert rc == 0
        sel = np.array([(slots >> l) & 1 for l in range(32)], bool)
        assert np.array_equal(out[0][sel], out[2][sel]) and np.array_equal(out[1][sel], out[3][sel]), hex(slots)
        assert not out[0][~sel].any() and not out[1][~sel].any()


@pytest.mark.parametrize("dim,bits,k,stats", [(128, 4, 10, False), (64, 4, 10, True), (32, 1, 140, False)])
def test_result_set_follows_the_reference_heap_where_distinct_ids_tie(k1, k3, oracle, dim, bits, k, stats):
    """Forty vectors stored under fifteen ids each: distinct ids of bit-equal distance meet at the result set's eviction boundary
    all the time, and which of them BoundedMaxHeap lets go is a matter of its heap layout (search/rabitq_search.hpp:26-35); an
    ascending list that evicts its last entry differs from it in a quarter of these rows."""
    fab = common.fabricate(600, dim, bits, seed=7 + dim, counts=(32, 32, 30, 12), layers=1)
    fab.raw[:] = fab.raw[np.arange(600) % 40]
    fab.norm_sq[:] = fab.norm_sq[np.arange(600) % 40]
    q = np.random.default_rng(3).standard_normal((16, dim)).astype(np.float32)
    ids, dists, over, _ = _search(k1, k3, oracle, dim, bits, fab.search_data, fab.search_data.shape[1], fab.nb_off, fab.raw, fab.norm_sq,
                                  fab.calibration, fab.max_level, fab.entry_point, fab.layers, q, k, want_stats=stats)
    assert over == 0
    oid, od, _ = oracle.search_batch(oracle.index_view(fab), q, k)
    assert np.array_equal(ids, oid) and np.array_equal(_bits(dists), _bits(od))   # entry for entry: sort_heap's order of equal distances too
