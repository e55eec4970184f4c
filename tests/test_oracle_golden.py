"""CPU: the oracle's C restatement against the committed golden vectors (produced by the unmodified
reference, tests/golden/make_golden.py) and -- where oracle/_ref is present -- against the compiled
reference itself on fresh random inputs.  This is what pins the oracle."""
import numpy as np
import pytest

import common
from common import GOLDEN, co


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def test_rotation_signs_self_check(oracle):
    # first 16 layer-0 signs for seed 42 are +++-+-+----+++++
    s = oracle.rotation_signs(128, 42)
    assert "".join("+" if x > 0 else "-" for x in s[0, :16]) == "+++-+-+----+++++"


@pytest.mark.parametrize("dim", [16, 20, 96, 128, 960])
def test_query_encoding_golden(oracle, dim):
    g = np.load(GOLDEN / "k1_golden.npz")
    lut, coeffs, rot = oracle.encode_queries(g[f"q_{dim}"], want_rotated=True)
    assert np.array_equal(lut, g[f"lut_{dim}"])
    assert np.array_equal(_bits(coeffs), _bits(g[f"coeffs_{dim}"]))
    # rotate_raw_vector (encoder/rabitq_encoder.hpp:81-86) = rotation + norm_factor scaling
    assert np.array_equal(_bits(rot), _bits(g[f"rot_{dim}"]))


@pytest.mark.parametrize("tag", ["128_1", "128_2", "128_4", "960_2", "16_4"])
def test_fastscan_golden(oracle, tag):
    g = np.load(GOLDEN / "k2_golden.npz")
    dim, bits = map(int, tag.split("_"))
    D = 1 << (dim - 1).bit_length()
    lay = co.nb_layout(D, bits)
    blocks, lut, coeffs, cal = g[f"blocks_{tag}"], g[f"lut_{tag}"], g[f"coeffs_{tag}"], g[f"calib_{tag}"]
    for v in range(blocks.shape[0]):
        rec = blocks[v]
        f = lambda name, t, cnt: rec[lay[name]:lay[name] + cnt].view(t)  # noqa: E731
        qi, count, dqp = int(g[f"qi_{tag}"][v]), int(g[f"count_{tag}"][v]), float(g[f"dqp_{tag}"][v])
        nbit, msb, msb2 = oracle.fastscan(D, bits, lut[qi], rec[:4 * D * bits])
        assert np.array_equal(nbit, g[f"nbit_{tag}"][v]) and np.array_equal(msb, g[f"msb_{tag}"][v])
        assert np.array_equal(msb2, g[f"msb2_{tag}"][v])
        params = np.array([*coeffs[qi], cal[0], cal[1], cal[2], cal[3 + v % 3]], np.float32)
        wpop = f("wpop", np.uint16, 64) if bits > 1 else None
        est, lower, msb_lower = oracle.convert(D, bits, params, nbit, msb, msb2, f("nop", np.float32, 128), f("ip_qo", np.float32, 128),
                                               f("ip_cp", np.float32, 128), f("pop", np.uint16, 64), wpop, count, dqp)
        for got, name in ((est, "est"), (lower, "lower"), (msb_lower, "msb_lower")):
            assert np.array_equal(_bits(got[:count]), _bits(g[f"{name}_{tag}"][v][:count])), (name, v, count)


@pytest.mark.parametrize("D", [16, 128, 1024])
def test_exact_distance_golden(oracle, D):
    g = np.load(GOLDEN / "l2_golden.npz")
    for i in range(6):
        assert _bits(oracle.dot(g[f"a_{D}"][i], g[f"b_{D}"][i])) == _bits(g[f"dot_{D}"][i])
        assert _bits(oracle.l2(g[f"a_{D}"][i], g[f"b_{D}"][i])) == _bits(g[f"l2_{D}"][i])


@pytest.mark.parametrize("bits", [1, 2, 4])
@pytest.mark.parametrize("k", [1, 10, 50])
def test_search_golden(oracle, bits, k):
    g = np.load(GOLDEN / "e2e_golden.npz")
    sf = co.SaveFile(GOLDEN / f"ref_n300_d24_b{bits}.bin")
    ids, dists, _ = oracle.search_batch(oracle.index_view(sf), g["queries"], k)
    gi, gd = common.sorted_rows(ids, dists)
    wi, wd = common.sorted_rows(g[f"ids_b{bits}_k{k}"], g[f"dists_b{bits}_k{k}"])
    assert np.array_equal(gi, wi) and np.array_equal(_bits(gd), _bits(wd))


def test_save_file_parser_rejects_garbage(tmp_path):
    p = tmp_path / "bad.bin"
    p.write_bytes(b"\0" * 1024)
    with pytest.raises(RuntimeError, match="Invalid magic"):
        co.SaveFile(p)


def test_fabricated_index_roundtrips_through_save_file(oracle, tmp_path):
    fab = common.fabricate(200, 40, 2, seed=5, layers=2, counts=(32, 20))
    sf = co.SaveFile(common.write_save_file(fab, tmp_path / "f.bin"))
    assert (sf.D, sf.B, sf.dim, sf.n, sf.max_level, sf.entry_point) == (fab.D, fab.B, fab.dim, fab.n, fab.max_level, fab.entry_point)
    assert np.array_equal(sf.search_data, fab.search_data) and np.array_equal(sf.raw, fab.raw)
    q = np.random.default_rng(0).standard_normal((10, 40)).astype(np.float32)
    a = oracle.search_batch(oracle.index_view(fab), q, 5)
    b = oracle.search_batch(oracle.index_view(sf), q, 5)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and a[2] == b[2]


# ---- live against the compiled reference (present in this container and shipped to the GPU box) ------
needs_ref = pytest.mark.skipif(not co.have_ref(), reason="oracle/_ref (compiled reference) not present")


@needs_ref
@pytest.mark.parametrize("dim", [16, 48, 96, 128, 200, 960, 2048])
def test_query_encoding_vs_reference(oracle, dim):
    q = np.random.default_rng(dim).standard_normal((32, dim)).astype(np.float32)
    q[0] = 0
    lut, coeffs = oracle.encode_queries(q)
    rlut, rco = oracle.ref_encode_queries(q)
    assert np.array_equal(lut, rlut) and np.array_equal(_bits(coeffs), _bits(rco))


@needs_ref
@pytest.mark.parametrize("dim,bits", [(128, 1), (128, 2), (128, 4), (64, 4), (1024, 2)])
def test_fastscan_and_epilogues_vs_reference(oracle, dim, bits):
    rng = np.random.default_rng(dim * bits)
    fab = common.fabricate(24, dim, bits, seed=3, counts=(32, 31, 25, 17, 16, 8, 3), degenerate=True, a=0.97, b=0.02)
    q = rng.standard_normal((3, dim)).astype(np.float32)
    lut, coeffs = oracle.encode_queries(q)
    lay, nb = fab.lay, fab.nb_off
    for v in range(fab.n):
        rec = fab.search_data[v]
        f = lambda name, t, cnt: rec[nb + lay[name]:nb + lay[name] + cnt].view(t)  # noqa: E731
        count = int(f("count", np.uint32, 4)[0])
        planes = rec[nb:nb + 4 * fab.D * bits]
        mine = oracle.fastscan(fab.D, bits, lut[v % 3], planes)
        ref = oracle.fastscan(fab.D, bits, lut[v % 3], planes, ref=True)
        for a, b in zip(mine, ref):
            assert np.array_equal(a, b)
        for dqp in (0.0, 3e-13, 1.5, 250.0):
            params = np.array([*coeffs[v % 3], fab.affine_a, fab.affine_b, fab.ip_qo_floor, 0.8], np.float32)
            wpop = f("wpop", np.uint16, 64) if bits > 1 else None
            args = (fab.D, bits, params, *mine, f("nop", np.float32, 128), f("ip_qo", np.float32, 128), f("ip_cp", np.float32, 128),
                    f("pop", np.uint16, 64), wpop, count, dqp)
            for a, b in zip(oracle.convert(*args), oracle.convert(*args, ref=True)):
                assert np.array_equal(_bits(a[:count]), _bits(b[:count]))


@needs_ref
@pytest.mark.parametrize("bits", [1, 4])
def test_search_vs_reference_module(oracle, bits):
    path = common.reference_index_file(5000, 64, bits)
    q = common.queries_for(64, 100)
    ref = co.ref_module().CPIndex(dim=64, bits=bits)
    ref.load(str(path))
    for k in (10, 0, 37):
        rid, rd = ref.search_batch(q, k)
        oid, od, _ = oracle.search_batch(oracle.index_view(co.SaveFile(path)), q, k)
        assert np.array_equal(rid, oid) and np.array_equal(_bits(rd), _bits(od))


@needs_ref
def test_search_vs_reference_module_where_distinct_ids_tie(oracle, tmp_path):
    """Sixty vectors stored twenty times each, index built by the unmodified reference: every result row holds distinct ids of
    bit-equal distance, and which of them survive an eviction is a matter of BoundedMaxHeap's layout (search/rabitq_search.hpp:
    26-35).  The restatement must return the reference's rows entry for entry -- this pins the heap the kernels are held to
    (tests/test_parity_gpu.py, tests/test_kernels_emulated.py: ..._where_distinct_ids_tie)."""
    rng = np.random.default_rng(5)
    base = np.tile(rng.standard_normal((60, 32)).astype(np.float32), (20, 1))
    rng.shuffle(base)
    path = tmp_path / "dup.bin"
    ref = co.build_reference_index(base, 4, path, threads=4)
    q = rng.standard_normal((300, 32)).astype(np.float32)
    view = oracle.index_view(co.SaveFile(path))
    for k in (10, 3, 25):
        rid, rd = ref.search_batch(q, k)
        oid, od, _ = oracle.search_batch(view, q, k)
        assert np.array_equal(rid, oid) and np.array_equal(_bits(rd), _bits(od))
        tied = sum(any(rd[r, a] == rd[r, b] and rid[r, a] != rid[r, b] for a in range(k) for b in range(a + 1, k)) for r in range(300))
        assert tied > 250   # the situation the test is about does occur


def test_exhaustive_restatement_against_the_reference_composition(oracle):
    """cpo_exhaustive_search against tests/golden/exhaustive_golden.npz, which was composed in Python from primitives
    executed by the unmodified reference (tests/golden/make_exhaustive_golden.py): integer sums, estimate bits, and the
    (k, k') results, on the committed reference-built 1-bit index."""
    g = np.load(common.GOLDEN / "exhaustive_golden.npz")
    sf = co.SaveFile(common.GOLDEN / "ref_n300_d24_b1.bin")
    view = oracle.index_view(sf)
    q = g["queries"]
    for i in range(len(q)):
        for k, kp in ((10, 100), (1, 1), (5, 32), (10, 300)):
            ids, dists, sums, est = oracle.exhaustive(view, sf, q[i], k, kp)
            assert np.array_equal(sums, g[f"sums_{i}"]), i
            assert np.array_equal(est.view(np.uint32), g[f"est_{i}"].view(np.uint32)), i
            wi, wd = g[f"ids_{i}_k{k}_kp{kp}"], g[f"dists_{i}_k{k}_kp{kp}"]
            m = int((wi >= 0).sum())
            assert len(ids) == m and np.array_equal(ids.astype(np.int64), wi[:m]), (i, k, kp)
            assert np.array_equal(np.asarray(dists, np.float32).view(np.uint32), wd[:m].view(np.uint32))
